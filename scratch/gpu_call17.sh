#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./scratch/micro/bulk_probe > gpurun_out/bulk_probe.txt 2>&1; echo "rc=$?" >> gpurun_out/bulk_probe.txt
timeout 600 python -m pytest tests -m gpu -q -x -k "duct_sst_residual or tfp or smoothbump" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_g3.log 2>&1
cat gpurun_out/bulk_probe.txt; tail -3 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/bench_g3.log | cut -c1-1100
