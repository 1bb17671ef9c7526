#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_g3.log 2>&1
tail -30 gpurun_out/pytest_gpu.log | cut -c1-300; tail -1 gpurun_out/bench_g3.log | cut -c1-400
