#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck_g3.log 2>&1; echo "rc=$?" >> gpurun_out/racecheck_g3.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_g3.log 2>&1
F3D_GRAD_TMA=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_g3_gradold.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gradients_tma --launch-skip 3 -c 1 -o gpurun_out/grad_full -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_grad.log 2>&1
tail -3 gpurun_out/racecheck_g3.log; tail -4 gpurun_out/pytest_gpu.log | cut -c1-250
for f in bench_g3 bench_g3_gradold; do tail -1 gpurun_out/$f.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f', d['ms_per_step'], d['roofline']['kernel_ms'], d['value'])"; done
grep -h "k_gradient" gpurun_out/launches.csv | tail -2 | cut -c1-260
