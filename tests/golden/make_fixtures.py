"""Generates tests/golden/{smoothbump,lfp,tfp}: the reference's own input files for BASELINE configs 1-3
(grids, layout, mapping, periodic, bc, control / fvscheme / flow) plus the shipped Tfp restart state, copied
verbatim from /root/reference/tests/<Case>/ (data files, not source code).  Run in the build container:

    python tests/golden/make_fixtures.py

Also writes tests/golden/oracle_vectors.npz: outputs of the CPU oracle on these cases, so that `-m "not gpu"`
tests pin the oracle build against drift (the vectors are the oracle's own, NOT reference outputs -- the reference
cannot be run here; parity with the Fortran build stays unpinned).

And tests/golden/smoothbump_reference_output.npz: the ONE reference output the tree ships for this path -- the flow field of the
reference's own SmoothBump run (tests/SmoothBump/time_directories/0010/process_0{0,1}.dat: Density, u, v, w, Pressure of the
interior cells, 16 significant digits), stored as arrays.  tests/test_oracle_kat.py uses it as a soft pin (entropy measure of
tests/SmoothBump/pp/entropy.py against the value in tests/Report.txt, and near-stationarity under the oracle's operator)."""
import importlib
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/tests"
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

CASES = {"smoothbump": "SmoothBump", "lfp": "Lfp", "tfp": "Tfp"}


def copy_case(name, ref):
    dst = os.path.join(HERE, name)
    shutil.rmtree(dst, ignore_errors=True)
    src = os.path.join(REF, ref, "system")
    for sub in ("", "mesh/gridfiles", "mesh/bc", "mesh/layout"):
        os.makedirs(os.path.join(dst, "system", sub), exist_ok=True)
    for f in ("control.md", "fvscheme.md", "flow.md"):
        shutil.copy(os.path.join(src, f), os.path.join(dst, "system", f))
    for f in os.listdir(os.path.join(src, "mesh/gridfiles")):
        shutil.copy(os.path.join(src, "mesh/gridfiles", f), os.path.join(dst, "system/mesh/gridfiles", f))
    for f in os.listdir(os.path.join(src, "mesh/bc")):
        shutil.copy(os.path.join(src, "mesh/bc", f), os.path.join(dst, "system/mesh/bc", f))
    for f in ("layout.md", "mapping.txt", "periodic.txt"):
        p = os.path.join(src, "mesh/layout", f)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(dst, "system/mesh/layout", f))
    if name == "tfp":   # shipped restart state = start state of the case (control.md 'Restart level 2')
        os.makedirs(os.path.join(dst, "restart"), exist_ok=True)
        for b in range(2):
            shutil.copy(os.path.join(REF, ref, "time_directories/0002/process_%02d.dat" % b), os.path.join(dst, "restart"))


def main():
    for name, ref in CASES.items():
        copy_case(name, ref)
    case_mod = importlib.import_module("fest3d_b200.case")
    import fixtures
    import oracle_py
    out = {}
    settings = {
        "smoothbump": dict(scheme=dict(time_step_accuracy="RK4"), control=dict(CFL=0.5)),
        "lfp": dict(scheme=dict(scheme_name="slau", interpolant="muscl", time_step_accuracy="RK4"), control=dict(CFL=0.5)),
        "tfp": dict(scheme=dict(scheme_name="ausmUP", interpolant="muscl", time_step_accuracy="RK4"), control=dict(CFL=0.5)),
    }
    for name in CASES:
        blocks = fixtures.load(case_mod, os.path.join(HERE, name), **settings[name])
        w = oracle_py.OracleWorld(blocks)
        err, res = w.residual(1)
        assert err == 0
        hist = []
        for it in range(1, 11):
            err, r = w.step(it)
            assert err == 0
            hist.append(r)
        out[name + "_hist"] = np.array(hist)
        for b in range(len(blocks)):
            out["%s_res%d" % (name, b)] = res[b]
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **out)
    # the reference's own converged SmoothBump field (a reference OUTPUT)
    blocks = fixtures.load(case_mod, os.path.join(HERE, "smoothbump"))
    ref_out = {}
    for b, blk in enumerate(blocks):
        case_mod.read_tecplot_state(os.path.join(REF, "SmoothBump/time_directories/0010/process_%02d.dat" % b), blk)
        ref_out["q%d" % b] = blk.qp[:, 3:3 + blk.kmx - 1, 3:3 + blk.jmx - 1, 3:3 + blk.imx - 1].copy()
    np.savez_compressed(os.path.join(HERE, "smoothbump_reference_output.npz"), **ref_out)
    print("fixtures written:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
