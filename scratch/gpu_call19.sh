#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_g3.log 2>&1; echo "rc=$?" >> gpurun_out/smoke_g3.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_g3.log 2>&1
tail -2 gpurun_out/smoke_g3.log; tail -1 gpurun_out/bench_g3.log | cut -c1-1100
