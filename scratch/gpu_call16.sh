#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python scratch/walldist_bench.py 256 > gpurun_out/walldist_256.json 2> gpurun_out/walldist_256.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_g3.log 2>&1
tail -4 gpurun_out/pytest_gpu.log | cut -c1-300; cat gpurun_out/walldist_256.json; tail -2 gpurun_out/walldist_256.err; tail -1 gpurun_out/bench_g3.log | cut -c1-260
