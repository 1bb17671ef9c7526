#!/bin/bash
# BASELINE.json's other synthetic config (WENO + AUSM+ + SST at 256^3) measured once; memcheck of the smoke case and of the new set-up / checkpoint paths
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --interpolant weno --scheme ausmP > gpurun_out/bench_weno_ausmP.log 2>&1
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "geometry_on_device or checkpoint or restart" > gpurun_out/memcheck_new.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/memcheck_new.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck_smoke.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/memcheck_smoke.log
tail -1 gpurun_out/bench_weno_ausmP.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('weno+ausmP', d['ms_per_step'], d['roofline']['kernel_ms'], d['value'])"
tail -4 gpurun_out/memcheck_new.log | cut -c1-200; tail -4 gpurun_out/memcheck_smoke.log | cut -c1-200
