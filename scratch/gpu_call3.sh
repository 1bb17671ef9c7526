#!/bin/bash
# generation-3 evidence: full ncu capture of one k_sweep3 launch at 256^3 (source-level stall sampling included)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep3 --launch-skip 3 -c 1 -o gpurun_out/sweep3_full -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full3.log 2>&1
tail -3 gpurun_out/ncu_full3.log; ls -la gpurun_out/
