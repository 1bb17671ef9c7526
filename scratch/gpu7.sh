#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/g7_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/g7_smoke.txt; tail -2 gpurun_out/g7_smoke.txt
timeout 200 python scratch/small_case_timing.py > gpurun_out/g7_small.txt 2>&1; cat gpurun_out/g7_small.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/g7_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/g7_pytest.txt; tail -22 gpurun_out/g7_pytest.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/g7_bench.json 2> gpurun_out/g7_bench.err; cut -c1-250 gpurun_out/g7_bench.json; tail -3 gpurun_out/g7_bench.err
