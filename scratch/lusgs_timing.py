"""Cost of one implicit (LU-SGS) iteration at the benchmark block size, against the explicit single-stage step, same case."""
import sys, time, importlib
import numpy as np
sys.path.insert(0, ".")
import torch
syn = importlib.import_module("fest3d_b200.synthetic")
solver = importlib.import_module("fest3d_b200.solver")
for n in (128, 256):
    for ta, turb in (("none", "sst"), ("implicit", "sst"), ("implicit", "none")):
        blocks = syn.make_duct_blocks(n, turbulence=turb, time_step_accuracy=ta, CFL=0.5 if ta == "none" else 50.0)
        s = solver.Solver(blocks)
        s.iterate(3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        k = 5
        h = s.iterate(k)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / k
        print("n=%d %-8s %-4s: %.2f ms per iteration, %d launches per iteration, %.3f G cell-updates/s" % (n, ta, turb, dt * 1e3, s.blocks[0].launch_count() // 8, n ** 3 / dt / 1e9), flush=True)
        s.close()
        del s, blocks
