#!/bin/bash
# GPU call 1 of round 2: TMEM scratch probe, full GPU parity suite with the round-1 kernels (new parity-hole tests included), baseline bench
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/g1_smi.txt 2>&1
./scratch/micro/tmem_scratch_probe > gpurun_out/g1_tmem_probe.txt 2>&1; echo "tmem rc=$?" >> gpurun_out/g1_tmem_probe.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/g1_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/g1_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/g1_bench.json 2> gpurun_out/g1_bench.err
tail -5 gpurun_out/g1_tmem_probe.txt gpurun_out/g1_pytest.txt; cat gpurun_out/g1_bench.json | cut -c1-600
