#!/bin/bash
# GPU call 3: where does the generation-4 kernel spend its time?  per-warp phase split + ncu full capture
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
F3D_LIB=$PWD/fest-3d_b200/libfest3d_gpu_pt.so timeout 300 python scratch/run_steps.py --steps 3 > gpurun_out/g3_phase.txt 2>&1
cat gpurun_out/g3_phase.txt | tail -20
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -o gpurun_out/g3_fused python scratch/run_steps.py --steps 2 > gpurun_out/g3_ncu.log 2>&1
tail -3 gpurun_out/g3_ncu.log; ls -la gpurun_out/*.ncu-rep
