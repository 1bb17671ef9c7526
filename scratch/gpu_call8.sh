#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_g3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gradients --launch-skip 3 -c 1 -o gpurun_out/grad_full -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_grad.log 2>&1
tail -1 gpurun_out/bench_g3.log | cut -c1-1100
