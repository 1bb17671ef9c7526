#!/bin/bash
# WENO / WENO-NM non-linear weights as ratios (products of the other two denominators instead of three reciprocals): parity suite + the WENO + AUSM+ bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --interpolant weno --scheme ausmP > gpurun_out/bench_weno_ausmP_ratio.log 2>&1
tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
tail -1 gpurun_out/bench_weno_ausmP_ratio.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('weno ratio', d['ms_per_step'], d['roofline']['kernel_ms'], d['value'])"
