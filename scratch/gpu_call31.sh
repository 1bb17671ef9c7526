#!/bin/bash
# exact power-of-two folding in the viscous face + AUSM splitting as integer-tested selects + line-aligned k_gradients warps: parity, A/B benches, ncu captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_cur.log 2>&1
for v in gradunal flxref; do F3D_LIB=$PWD/scratch/libfest3d_gpu_$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$v.log 2>&1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gradient_bc --launch-skip 3 -c 1 -o gpurun_out/gradbc_full -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_gradbc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep3 --launch-skip 3 -c 1 -o gpurun_out/sweep3_full -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full3.log 2>&1
tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
for f in bench_cur bench_gradunal bench_flxref; do tail -1 gpurun_out/$f.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f', d['ms_per_step'], d['roofline']['kernel_ms'], d['value'])"; done
grep -h "k_gradient" gpurun_out/launches.csv | tail -2 | cut -c1-60
