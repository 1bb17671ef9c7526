#!/bin/bash
# asynchronous checkpoint / bitwise restart tests + the whole GPU parity suite + bench of the committed kernel set
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.log 2>&1
tail -5 gpurun_out/pytest_gpu.log | cut -c1-400
tail -1 gpurun_out/bench_final.log | cut -c1-400
