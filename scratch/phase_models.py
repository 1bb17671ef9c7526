"""Per-warp work / wait split of the default sweep (F3D_PHASE_TIMING build) for several models at 128^3."""
import sys, importlib, ctypes
sys.path.insert(0, ".")
import torch
syn = importlib.import_module("fest3d_b200.synthetic")
solver = importlib.import_module("fest3d_b200.solver")
capi = importlib.import_module("fest3d_b200.capi")
L = capi.lib()
for name, kw in (("sst", dict(turbulence="sst")), ("sst + bc", dict(turbulence="sst", transition="bc")), ("sa", dict(turbulence="sa")),
                 ("kkl", dict(turbulence="kkl")), ("sst + lctm2015", dict(turbulence="sst", transition="lctm2015"))):
    blocks = syn.make_duct_blocks(128, time_step_accuracy="none", CFL=0.5, **kw)
    s = solver.Solver(blocks)
    s.iterate(3)
    torch.cuda.synchronize()
    print("==", name, flush=True)
    L.fest3d_gpu_phase_dump3()      # totals so far (cumulative over models: differences matter)
    sys.stdout.flush()
    s.close()
