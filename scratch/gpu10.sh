#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "async or smoothbump_two or four_blocks" > gpurun_out/g10_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/g10_pytest.txt; tail -4 gpurun_out/g10_pytest.txt
( time timeout 600 python bench.py ) > gpurun_out/g10_bench.json 2> gpurun_out/g10_bench.err; cut -c1-250 gpurun_out/g10_bench.json; tail -5 gpurun_out/g10_bench.err

