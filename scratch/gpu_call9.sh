#!/bin/bash
# round-1 evidence of the generation-3 kernel: tests, bench (with the CPU leg), launch list, full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_256.log 2>&1
F3D_SWEEP_GEN=2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_256_gen2.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_reference.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep3 --launch-skip 3 -c 1 -o gpurun_out/sweep3_full -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gradients --launch-skip 3 -c 1 -o gpurun_out/grad_full -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_grad.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/bench_256.log | cut -c1-1500; tail -1 gpurun_out/bench_256_gen2.log | cut -c1-300; tail -1 gpurun_out/bench_reference.log | cut -c1-600
