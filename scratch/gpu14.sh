#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/g14_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/g14_pytest.txt; tail -6 gpurun_out/g14_pytest.txt | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g14_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/g14_smoke.txt; tail -4 gpurun_out/g14_smoke.txt
( time timeout 600 python bench.py ) > gpurun_out/g14_bench.json 2> gpurun_out/g14_bench.err; cut -c1-250 gpurun_out/g14_bench.json; tail -4 gpurun_out/g14_bench.err
