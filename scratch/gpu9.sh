#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
F3D_LIB=$PWD/fest-3d_b200/libfest3d_gpu_pt.so timeout 300 python scratch/run_steps.py --steps 3 > gpurun_out/g9_phase3.txt 2>&1; tail -17 gpurun_out/g9_phase3.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/g9_bench.json 2> gpurun_out/g9_bench.err; cut -c1-250 gpurun_out/g9_bench.json
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --interpolant weno --scheme ausmP --mode residual > gpurun_out/g9_bench_weno_res.json 2> gpurun_out/g9_bench_weno_res.err; cut -c1-250 gpurun_out/g9_bench_weno_res.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sweep3|k_gradients" -s 2 -c 2 -o gpurun_out/g9_staged python scratch/run_steps.py --steps 2 > gpurun_out/g9_ncu.log 2>&1; tail -2 gpurun_out/g9_ncu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/g9_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1; tail -3 gpurun_out/g9_launches.csv | cut -c1-200
