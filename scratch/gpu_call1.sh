#!/bin/bash
# one gpurun call: GPU parity tests, bench line, ncu launch list, ncu full capture of the top kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_256.log 2>&1
timeout 300 python bench.py --cells 128 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_128.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep --launch-skip 3 -c 1 -o gpurun_out/sweep_full -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gradients --launch-skip 3 -c 1 -o gpurun_out/grad_full -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_grad.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_256.log | tail -2; tail -1 gpurun_out/bench_128.log
