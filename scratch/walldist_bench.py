#!/usr/bin/env python
"""Device wall distance at the bench size: 256^3-cell duct, four no-slip walls (wall_dist.f90:84-131 brute force)."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
syn = importlib.import_module("fest3d_b200.synthetic")
geo = importlib.import_module("fest3d_b200.geometry")
solver = importlib.import_module("fest3d_b200.solver")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
blocks = syn.make_duct_blocks(n, turbulence="sst")
blk = blocks[0]
o = 2
I, J, K = slice(1 + o, blk.imx + o + 1), slice(1 + o, blk.jmx + o + 1), slice(1 + o, blk.kmx + o + 1)
N = blk.nodes
wall = np.concatenate([N[K, 1 + o, I].reshape(-1, 3), N[K, blk.jmx + o, I].reshape(-1, 3), N[1 + o, J, I].reshape(-1, 3), N[blk.kmx + o, J, I].reshape(-1, 3)])
s = solver.Solver(blocks)
t0 = time.perf_counter()
d, ms = s.blocks[0].find_wall_dist(wall, want_time=True)
wall_s = time.perf_counter() - t0
pairs = float(N.shape[0] * N.shape[1] * N.shape[2]) * len(wall)
ref = blk.dist   # analytic distance to the four walls used by the synthetic duct
K5, J5, I5 = slice(3, 3 + blk.kmx - 1), slice(3, 3 + blk.jmx - 1), slice(3, 3 + blk.imx - 1)
print(json.dumps({"cells": n ** 3, "nodes": int(pairs / len(wall)), "wall_nodes": int(len(wall)), "pairs": pairs, "kernel_ms": ms,
                  "pairs_per_s": pairs / (ms * 1e-3), "call_s_incl_copies": wall_s,
                  "max_abs_diff_to_analytic_wall_distance_over_h": float(np.abs(d[K5, J5, I5] - ref[K5, J5, I5]).max() * n)}))
s.close()
