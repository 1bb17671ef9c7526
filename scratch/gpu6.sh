#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 400 python scratch/cd_march.py lfp 200000 1.5 ausm RK4 > gpurun_out/g6_cd_lfp.txt 2>&1; tail -8 gpurun_out/g6_cd_lfp.txt
timeout 300 python scratch/cd_march.py tfp 100000 1.0 ausm RK4 > gpurun_out/g6_cd_tfp.txt 2>&1; tail -8 gpurun_out/g6_cd_tfp.txt
