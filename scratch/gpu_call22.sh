#!/bin/bash
# tensor-map TMA staging: probe, then the experimental library (F3D_LIB) through smoke, the parity suite and the bench, A/B with the default
mkdir -p gpurun_out
timeout 120 ./scratch/micro/tma4d_probe > gpurun_out/tma4d_probe.txt 2>&1; echo "rc=$?" >> gpurun_out/tma4d_probe.txt
export BULK=$PWD/scratch/libfest3d_gpu_bulk.so
F3D_LIB=$BULK timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_bulk.log 2>&1; echo "rc=$?" >> gpurun_out/smoke_bulk.log
F3D_LIB=$BULK timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_bulk.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_bulk.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_g3.log 2>&1
F3D_LIB=$BULK timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_bulk.log 2>&1
F3D_LIB=$BULK timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep3 --launch-skip 3 -c 1 -o gpurun_out/sweep3_bulk -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bulk.log 2>&1
cat gpurun_out/tma4d_probe.txt; tail -2 gpurun_out/smoke_bulk.log; tail -4 gpurun_out/pytest_bulk.log | cut -c1-300
tail -1 gpurun_out/bench_g3.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('default', d['ms_per_step'], d['roofline']['kernel_ms'])"
tail -1 gpurun_out/bench_bulk.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tma', d['ms_per_step'], d['roofline']['kernel_ms'])"
