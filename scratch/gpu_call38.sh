#!/bin/bash
mkdir -p gpurun_out
timeout 75 python scratch/smoothbump_march.py 5000 60000 > gpurun_out/smoothbump_march.txt 2>&1
tail -4 gpurun_out/smoothbump_march.txt
