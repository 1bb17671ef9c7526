#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/g8_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/g8_pytest.txt; tail -6 gpurun_out/g8_pytest.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/g8_bench.json 2> gpurun_out/g8_bench.err; cut -c1-250 gpurun_out/g8_bench.json; tail -3 gpurun_out/g8_bench.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --interpolant weno --scheme ausmP --mode residual > gpurun_out/g8_bench_weno_res.json 2> gpurun_out/g8_bench_weno_res.err; cut -c1-250 gpurun_out/g8_bench_weno_res.json
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --interpolant weno --scheme ausmP > gpurun_out/g8_bench_weno.json 2> gpurun_out/g8_bench_weno.err; cut -c1-250 gpurun_out/g8_bench_weno.json
