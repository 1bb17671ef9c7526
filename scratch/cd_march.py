#!/usr/bin/env python
"""March the reference's flat-plate cases on the device to a steady drag coefficient (explicit, local time step) and print C_d
as the reference's post-processing computes it (tests/surface_drag.py): usage cd_march.py lfp|tfp [iterations] [CFL] [scheme] [ta]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import fixtures, surface_drag
case = sys.argv[1]; n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 100000; cfl = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
scheme = sys.argv[4] if len(sys.argv) > 4 else "ausm"; ta = sys.argv[5] if len(sys.argv) > 5 else "RK4"
case_mod = importlib.import_module("fest3d_b200.case"); solver = importlib.import_module("fest3d_b200.solver")
blocks = fixtures.load(case_mod, os.path.join(ROOT, "tests", "golden", case), scheme=dict(scheme_name=scheme, interpolant="muscl", time_step_accuracy=ta), control=dict(CFL=cfl))
s = solver.Solver(blocks)
print("%s: %s + muscl, %s, CFL %g; report %.4e, script target %.3e" % (case, scheme, ta, cfl, surface_drag.CD_REPORT[case], surface_drag.CD_EXPECTED[case]), flush=True)
print("iter 0: C_d %.5e" % surface_drag.device_wall_drag(s, blocks, case), flush=True)
t0 = time.time(); done = 0; chunk = max(n_it // 20, 1)
while done < n_it:
    r = s.iterate(chunk - 1, want_norms=False) if chunk > 1 else None
    r = s.iterate(1)[0]; done += chunk
    cd = surface_drag.device_wall_drag(s, blocks, case)
    print("iter %7d: C_d %.5e (%.3f %% of the report)  mass residual %.3e  x-mom %.3e   %.1f s" % (done, cd, 100 * cd / surface_drag.CD_REPORT[case], r[1], r[2], time.time() - t0), flush=True)
s.close()
