#!/bin/bash
# the GPU soft-pin test on the reference's shipped SmoothBump output (3000 explicit iterations on two 48x48 blocks)
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "shipped_smoothbump" > gpurun_out/pytest_softpin.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_softpin.log
tail -6 gpurun_out/pytest_softpin.log | cut -c1-300
