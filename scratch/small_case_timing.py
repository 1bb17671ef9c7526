#!/usr/bin/env python
"""Launch-bound small case: SmoothBump (2 blocks of 48 x 48 x 1 cells), RK4: milliseconds per iteration with and without CUDA graphs
(F3D_GRAPHS=0), and bitwise equality of the two marches."""
import importlib, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import fixtures
if len(sys.argv) > 1 and sys.argv[1] == "child":
    case_mod = importlib.import_module("fest3d_b200.case"); solver = importlib.import_module("fest3d_b200.solver")
    blocks = fixtures.load(case_mod, os.path.join(ROOT, "tests", "golden", "smoothbump"), scheme=dict(time_step_accuracy="RK4"), control=dict(CFL=1.0))
    s = solver.Solver(blocks)
    s.iterate(50)
    s.blocks[0].sync(); t0 = time.perf_counter()
    h = s.iterate(2000)
    for b in s.blocks: b.sync()
    dt = time.perf_counter() - t0
    q = s.blocks[0].get_state()
    print("graphs=%s  %.4f ms per iteration  launches %d  checksum %.17g %.17g" % (os.environ.get("F3D_GRAPHS", "1"), 1e3 * dt / 2000, s.blocks[0].launch_count(), float(h[-1].sum()), float(q.sum())))
    s.close()
else:
    for g in ("1", "0"):
        print(subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, F3D_GRAPHS=g), capture_output=True, text=True).stdout.strip())
