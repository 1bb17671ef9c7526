#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "lctm or kkl" > gpurun_out/g13_lctm.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/g13_lctm.txt; tail -60 gpurun_out/g13_lctm.txt | cut -c1-300
