#!/bin/bash
# A/B: default build (per-thread cp.async) vs the bulk-copy staging build (copies issued by all threads)
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_g3.log 2>&1
F3D_LIB=$PWD/scratch/libfest3d_gpu_bulk.so timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_bulk.log 2>&1; echo "rc=$?" >> gpurun_out/smoke_bulk.log
F3D_LIB=$PWD/scratch/libfest3d_gpu_bulk.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_bulk.log 2>&1
F3D_LIB=$PWD/scratch/libfest3d_gpu_bulk.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep3 --launch-skip 3 -c 1 -o gpurun_out/sweep3_bulk -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bulk.log 2>&1
tail -1 gpurun_out/bench_g3.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('default', d['ms_per_step'], d['roofline']['kernel_ms'])"
tail -2 gpurun_out/smoke_bulk.log
tail -1 gpurun_out/bench_bulk.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bulk', d['ms_per_step'], d['roofline']['kernel_ms'])"
