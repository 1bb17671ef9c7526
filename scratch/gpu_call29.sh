#!/bin/bash
# device geometry parity (new tests), integer-range Koren psi: full GPU parity suite, bench A/B against the reference-form psi from the same sources, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_psi.log 2>&1
F3D_LIB=$PWD/scratch/libfest3d_gpu_korenref.so timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_psiref.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
for f in bench_psi bench_psiref; do tail -1 gpurun_out/$f.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f', d['ms_per_step'], d['roofline']['kernel_ms'], d['value'])"; done
grep -h "k_gradient" gpurun_out/launches.csv | tail -2 | cut -c1-200
