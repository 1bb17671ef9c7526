#!/bin/bash
# GPU call 2: first run of the generation-4 fused kernel (gradients in the tile pass, TMEM hand-overs)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
# quick smoke first with a hard timeout (a hang must not eat the budget)
timeout 120 python __graft_entry__.py smoke > gpurun_out/g2_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/g2_smoke.txt
tail -3 gpurun_out/g2_smoke.txt
timeout 300 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/g2_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/g2_memcheck.txt
tail -5 gpurun_out/g2_memcheck.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g2_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/g2_pytest.txt
tail -15 gpurun_out/g2_pytest.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/g2_bench.json 2> gpurun_out/g2_bench.err
cut -c1-400 gpurun_out/g2_bench.json; tail -3 gpurun_out/g2_bench.err
F3D_LIB=$PWD/fest-3d_b200/libfest3d_gpu_g3.so timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/g2_bench_g3.json 2> gpurun_out/g2_bench_g3.err
cut -c1-400 gpurun_out/g2_bench_g3.json
