#!/bin/bash
# retry a gpurun call while the pod answers "transient" (nothing charged); usage: gpurun_retry.sh <timeout> <script> [gpus]
for n in 1 2 3 4 5 6 7 8 9 10 11 12; do
  if [ -n "$3" ]; then out=$(/usr/local/graft/bin/gpurun --gpus $3 --timeout $1 -- "bash $2" 2>&1); else out=$(/usr/local/graft/bin/gpurun --timeout $1 -- "bash $2" 2>&1); fi
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  echo "$out" | tail -12 | cut -c1-700; exit 0
done
echo "gave up: pod busy"
