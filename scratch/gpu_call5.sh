#!/bin/bash
mkdir -p gpurun_out
./scratch/micro/rcp_accuracy > gpurun_out/rcp_accuracy.txt 2>&1; echo "rc=$?" >> gpurun_out/rcp_accuracy.txt
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_g3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep3 --launch-skip 3 -c 1 -o gpurun_out/sweep3_full -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full3.log 2>&1
cat gpurun_out/rcp_accuracy.txt; tail -4 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/bench_g3.log | cut -c1-1100
