#!/bin/bash
# L2 prefetch of the staged planes (UTMAPF two planes ahead) + the K rows' q(k+3) fetched after their reconstruction: racecheck, parity, A/B benches, barrier-wait diagnostic
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool racecheck --racecheck-report analysis python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck_g3.log 2>&1; echo "rc=$?" >> gpurun_out/racecheck_g3.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_pf.log 2>&1
for v in nol2pf q2same pfmet; do F3D_LIB=$PWD/scratch/libfest3d_gpu_$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$v.log 2>&1; done
F3D_LIB=$PWD/scratch/libfest3d_gpu_arrival.so timeout 300 python scratch/arrival_timing.py 256 > gpurun_out/arrival.txt 2>&1
tail -3 gpurun_out/racecheck_g3.log; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
for f in bench_pf bench_nol2pf bench_q2same bench_pfmet; do tail -1 gpurun_out/$f.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f', d['ms_per_step'], d['roofline']['kernel_ms'], d['value'])"; done
tail -16 gpurun_out/arrival.txt
