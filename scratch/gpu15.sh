#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s -k "lusgs or unsupported_is" > gpurun_out/g15_lusgs.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/g15_lusgs.txt; tail -60 gpurun_out/g15_lusgs.txt | cut -c1-400
