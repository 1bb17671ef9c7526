#!/usr/bin/env python
"""A few steps of the bench workload (256^3 MUSCL + AUSM + SST by default) without the bench's extra legs: for ncu captures and the
phase-timing build (F3D_LIB=...libfest3d_gpu_pt.so prints the per-warp phase split)."""
import argparse, ctypes, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=256)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--scheme", default="ausm"); ap.add_argument("--interpolant", default="muscl"); ap.add_argument("--ta", default="none")
args = ap.parse_args()
syn = importlib.import_module("fest3d_b200.synthetic"); solver = importlib.import_module("fest3d_b200.solver"); capi = importlib.import_module("fest3d_b200.capi")
blocks = syn.make_duct_blocks(args.cells, scheme_name=args.scheme, interpolant=args.interpolant, turbulence="sst", time_step_accuracy=args.ta, CFL=0.5)
s = solver.Solver(blocks)
s.iterate(1, want_norms=False)
L = capi.lib()
if hasattr(L, "fest3d_gpu_phase_dump"):
    L.fest3d_gpu_phase_dump()   # discard the warm-up
r = s.iterate(args.steps)
print("res_abs last:", r[-1])
if hasattr(L, "fest3d_gpu_phase_dump"):
    L.fest3d_gpu_phase_dump()
if hasattr(L, "fest3d_gpu_phase_dump3"):
    L.fest3d_gpu_phase_dump3()
s.close()
