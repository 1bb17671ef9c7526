#!/bin/bash
# generation-3 bring-up: racecheck of the smoke case, GPU parity tests, A/B bench of generation 3 vs 2 from the same build
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/smi.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_g3.log 2>&1; echo "rc=$?" >> gpurun_out/smoke_g3.log
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck_g3.log 2>&1; echo "rc=$?" >> gpurun_out/racecheck_g3.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_g3.log 2>&1
F3D_SWEEP_GEN=2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_g2.log 2>&1
tail -3 gpurun_out/smoke_g3.log; tail -5 gpurun_out/racecheck_g3.log; tail -5 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/bench_g3.log | cut -c1-900; tail -1 gpurun_out/bench_g2.log | cut -c1-900
