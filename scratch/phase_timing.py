"""Development aid: build with `make EXTRA=-DF3D_PHASE_TIMING` and print the per-phase clock shares of a main warp."""
import ctypes, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
syn = importlib.import_module("fest-3d_b200.synthetic"); solver = importlib.import_module("fest-3d_b200.solver"); capi = importlib.import_module("fest-3d_b200.capi")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
blocks = syn.make_duct_blocks(n, turbulence="sst", time_step_accuracy="none", CFL=0.5)
s = solver.Solver(blocks)
s.iterate(3, want_norms=False)
capi.lib().fest3d_gpu_phase_dump()
s.iterate(5, want_norms=False)
capi.lib().fest3d_gpu_phase_dump()
