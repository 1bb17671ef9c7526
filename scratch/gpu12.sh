#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "kkl" > gpurun_out/g12_kkl.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/g12_kkl.txt; tail -40 gpurun_out/g12_kkl.txt
