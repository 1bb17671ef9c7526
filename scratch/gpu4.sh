#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/g4_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/g4_smoke.txt; tail -2 gpurun_out/g4_smoke.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g4_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/g4_pytest.txt; tail -6 gpurun_out/g4_pytest.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/g4_bench.json 2> gpurun_out/g4_bench.err; cut -c1-330 gpurun_out/g4_bench.json
F3D_LIB=$PWD/fest-3d_b200/libfest3d_gpu_pt.so timeout 300 python scratch/run_steps.py --steps 3 > gpurun_out/g4_phase.txt 2>&1; tail -17 gpurun_out/g4_phase.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -o gpurun_out/g4_fused python scratch/run_steps.py --steps 2 > gpurun_out/g4_ncu.log 2>&1; tail -2 gpurun_out/g4_ncu.log
