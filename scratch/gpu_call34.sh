#!/bin/bash
# WENO / WENO-NM / PPM divisions through the fast reciprocal + a dedicated WENO + AUSM+ instantiation: parity suite, A/B at 256^3 against IEEE divisions, headline bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --interpolant weno --scheme ausmP > gpurun_out/bench_weno_ausmP_fast.log 2>&1
F3D_LIB=$PWD/scratch/libfest3d_gpu_wenoieee.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --interpolant weno --scheme ausmP > gpurun_out/bench_weno_ausmP_ieee.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_cur.log 2>&1
tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
for f in bench_weno_ausmP_fast bench_weno_ausmP_ieee bench_cur; do tail -1 gpurun_out/$f.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f', d['ms_per_step'], d['roofline']['kernel_ms'], d['value'])"; done
