#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "duct_sst_residual or tfp or lfp or sa_residual or time_integrators" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_split.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
F3D_GRAD_SPLIT=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_nosplit.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch2.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_g3.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; grep -h "k_gradients" gpurun_out/launches_split.csv | tail -2 | cut -c1-200; grep -h "k_gradients" gpurun_out/launches_nosplit.csv | tail -2 | cut -c1-200; tail -1 gpurun_out/bench_g3.log | cut -c1-300
