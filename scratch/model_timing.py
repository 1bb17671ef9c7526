"""Per-step cost of every model of the staged path at the benchmark block size (256^3, muscl + ausm, single-stage update)."""
import sys, time, importlib
sys.path.insert(0, ".")
import torch
syn = importlib.import_module("fest3d_b200.synthetic")
solver = importlib.import_module("fest3d_b200.solver")
import os
n = 256
only = os.environ.get("MODELS")
for name, kw in (("none (laminar)", dict(turbulence="none")), ("sa", dict(turbulence="sa")), ("sst", dict(turbulence="sst")), ("sst + bc", dict(turbulence="sst", transition="bc")),
                 ("kkl", dict(turbulence="kkl")), ("sst + lctm2015", dict(turbulence="sst", transition="lctm2015"))):
    if only and name not in only.split(","):
        continue
    blocks = syn.make_duct_blocks(n, time_step_accuracy="none", CFL=0.5, **kw)
    s = solver.Solver(blocks)
    s.iterate(3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    k = 20
    s.iterate(k)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / k
    print("%-16s n_var %d: %.2f ms per step, %.3f G cell-updates/s" % (name, blocks[0].n_var, dt * 1e3, n ** 3 / dt / 1e9), flush=True)
    s.close()
    del s, blocks
