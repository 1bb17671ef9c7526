#!/bin/bash
# round-1 evidence of the TMA-staged generation-3 kernel: racecheck, tests, bench (with the CPU leg), launch list, full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/smi.txt
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck_g3.log 2>&1; echo "rc=$?" >> gpurun_out/racecheck_g3.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 40 --warmup 3 > gpurun_out/bench_256.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep3 --launch-skip 3 -c 1 -o gpurun_out/sweep3_full -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full3.log 2>&1
tail -3 gpurun_out/racecheck_g3.log; tail -4 gpurun_out/pytest_gpu.log | cut -c1-200; tail -1 gpurun_out/bench_256.log | cut -c1-1500
