#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_g3.log 2>&1; echo "rc=$?" >> gpurun_out/smoke_g3.log
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck_g3.log 2>&1; echo "rc=$?" >> gpurun_out/racecheck_g3.log
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_g3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep3 --launch-skip 3 -c 1 -o gpurun_out/sweep3_full -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full3.log 2>&1
tail -2 gpurun_out/smoke_g3.log; tail -3 gpurun_out/racecheck_g3.log; tail -4 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/bench_g3.log | cut -c1-1100
