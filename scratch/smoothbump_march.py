"""Evidence run: the CUDA path marched from the reference's shipped SmoothBump output (tests/golden/smoothbump_reference_output.npz);
prints residual norms and the entropy measure of tests/SmoothBump/pp/entropy.py (tests/Report.txt reports 7.883e-07)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
case_mod = importlib.import_module("fest3d_b200.case"); solver = importlib.import_module("fest3d_b200.solver")
import fixtures
G = os.path.join(ROOT, "tests", "golden")
blocks = fixtures.load(case_mod, os.path.join(G, "smoothbump"), scheme=dict(time_step_accuracy="RK4"), control=dict(CFL=1.0))
ref = np.load(os.path.join(G, "smoothbump_reference_output.npz"))
for b, blk in enumerate(blocks):
    blk.qp[:, 3:3 + blk.kmx - 1, 3:3 + blk.jmx - 1, 3:3 + blk.imx - 1] = ref["q%d" % b]
s = solver.Solver(blocks)
def ds():
    e2 = v = 0.0
    for gb, blk in zip(s.blocks, blocks):
        q = gb.get_state(); nk, nj, ni = blk.kmx - 1, blk.jmx - 1, blk.imx - 1
        rho, p = q[0, 3:3 + nk, 3:3 + nj, 3:3 + ni], q[4, 3:3 + nk, 3:3 + nj, 3:3 + ni]
        V = blk.cells[3:3 + nk, 3:3 + nj, 3:3 + ni, 0]; si = blk.flow.pressure_inf / blk.flow.density_inf ** 1.4
        e2 += ((((p / rho ** 1.4) - si) * V / si) ** 2).sum(); v += V.sum()
    return float(np.sqrt(e2 / v))
t0 = time.time(); n = 0
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
total = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
print("iter %7d  Ds %.6e" % (0, ds()), flush=True)
while n < total:
    s.iterate(chunk - 1, want_norms=False); r = s.iterate(1)[0]; n += chunk
    print("iter %7d  mass %.4e  x-mom %.4e  energy %.4e  Ds %.6e  (%.1f s)" % (n, r[1], r[2], r[5], ds(), time.time() - t0), flush=True)
