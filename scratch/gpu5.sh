#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/g5_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/g5_smoke.txt; tail -2 gpurun_out/g5_smoke.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/g5_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/g5_pytest.txt; tail -8 gpurun_out/g5_pytest.txt
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/g5_bench.json 2> gpurun_out/g5_bench.err; cut -c1-250 gpurun_out/g5_bench.json; tail -3 gpurun_out/g5_bench.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --gradients fused > gpurun_out/g5_bench_fused.json 2> gpurun_out/g5_bench_fused.err; cut -c1-250 gpurun_out/g5_bench_fused.json
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --interpolant weno --scheme ausmP --mode residual > gpurun_out/g5_bench_weno_res.json 2> gpurun_out/g5_bench_weno_res.err; cut -c1-250 gpurun_out/g5_bench_weno_res.json; tail -2 gpurun_out/g5_bench_weno_res.err
free -g | head -2
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --scaling strong > gpurun_out/g5_bench_strong1.json 2> gpurun_out/g5_bench_strong1.err; cut -c1-250 gpurun_out/g5_bench_strong1.json; tail -3 gpurun_out/g5_bench_strong1.err
