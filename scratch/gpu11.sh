#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "kkl" > gpurun_out/g11_kkl.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/g11_kkl.txt; tail -30 gpurun_out/g11_kkl.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "not kkl and not drag" > gpurun_out/g11_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/g11_pytest.txt; tail -4 gpurun_out/g11_pytest.txt
( time timeout 600 python bench.py ) > gpurun_out/g11_bench.json 2> gpurun_out/g11_bench.err; cut -c1-250 gpurun_out/g11_bench.json; tail -5 gpurun_out/g11_bench.err
