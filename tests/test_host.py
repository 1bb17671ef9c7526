"""CPU-side tests: the C-ABI library loads and exports every symbol include/fest3d_gpu.h declares (no compute without a
GPU), the config struct agrees between header and binding, the host stand-in reads the reference's case files, the
geometry obeys its invariants, and the oracle reproduces its committed golden vectors."""
import ctypes as C
import importlib
import os
import re

import numpy as np
import pytest

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_abi_exports_every_declared_symbol():
    capi = importlib.import_module("fest3d_b200.capi")
    hdr = open(os.path.join(ROOT, "include", "fest3d_gpu.h")).read()
    declared = sorted(set(re.findall(r"\b(fest3d_gpu_[a-z_0-9]+)\s*\(", hdr)))
    assert declared == sorted(capi.SYMBOLS)
    lib = capi.lib()
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert b"sm_100a" in lib.fest3d_gpu_version()


def test_no_device_is_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    capi = importlib.import_module("fest3d_b200.capi")
    lib = capi.lib()
    h = C.c_void_p()
    cfg = capi.Fest3dGpuConfig()
    assert lib.fest3d_gpu_create(C.byref(h), C.byref(cfg), 0) != 0
    assert not h.value


def test_config_struct_layout_matches_header(tmp_path):
    """The ctypes mirrors (product binding and oracle binding) against the C compiler's view of include/fest3d_gpu.h: size, the number
    of fixed-value slots, and the offsets of the first and last double fields."""
    import subprocess
    capi = importlib.import_module("fest3d_b200.capi")
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "fest3d_gpu.h"\n'
                   'int main(void) { printf("%zu %d %zu %zu %zu\\n", sizeof(Fest3dGpuConfig), F3D_NFIX, offsetof(Fest3dGpuConfig, gm),'
                   ' offsetof(Fest3dGpuConfig, tkl_inf), offsetof(Fest3dGpuConfig, fixed)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    size, nfix, off_gm, off_tkl, off_fixed = (int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split())
    cfg = capi.Fest3dGpuConfig
    assert C.sizeof(cfg) == size and capi.NFIX == nfix == 13
    assert cfg.gm.offset == off_gm and cfg.tkl_inf.offset == off_tkl and cfg.fixed.offset == off_fixed
    import oracle_py
    ocfg = oracle_py.OracleConfig
    assert C.sizeof(ocfg) == size and ocfg.gm.offset == off_gm and ocfg.tkl_inf.offset == off_tkl and ocfg.fixed.offset == off_fixed


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may import, link or execute it."""
    pk = os.path.join(ROOT, "fest-3d_b200")
    banned = ("oracle_py", "liboracle", "oracle_abi", "oracle_core", "import oracle", "from oracle", "oracle/")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                for b in banned:
                    assert b not in txt, (f, b)


def test_case_reader_smoothbump(case_mod):
    import fixtures
    blocks = fixtures.load(case_mod, os.path.join(GOLDEN, "smoothbump"))
    assert len(blocks) == 2
    b0, b1 = blocks
    assert (b0.imx, b0.jmx, b0.kmx) == (49, 49, 2)
    assert b0.bc_id == [-8, 1, -6, -6, -6, -6] and b1.bc_id == [0, -4, -6, -6, -6, -6]
    assert b0.scheme.scheme_name == "ausm" and b0.scheme.interpolant == "muscl" and b0.scheme.limiter == (0, 0, 0)
    assert b0.flow.mu_ref == 0.0 and abs(b0.flow.MInf - 170.14 / np.sqrt(1.4 * 101325.0 / 1.225)) < 1e-15
    assert b0.phi[1] == [48, 1] and b0.pdir[1] == [1, 1] and b0.otherface[1] == 1
    # the two blocks share the interface node plane exactly
    assert np.array_equal(b0.nodes[3:5, 3:52, 51], b1.nodes[3:5, 3:52, 3])


def test_case_reader_tfp_restart_and_walls(case_mod):
    import fixtures
    blocks = fixtures.load(case_mod, os.path.join(GOLDEN, "tfp"))
    b0, b1 = blocks
    assert b0.n_var == 7 and b1.bc_id[2] == -5
    f = b0.flow
    assert abs(f.tk_inf - 1.5 * (f.vel_mag * f.tu_inf / 100) ** 2) < 1e-18
    assert abs(f.tw_inf - f.density_inf * f.tk_inf / (f.mu_ref * f.mu_ratio_inf)) < 1e-9 * f.tw_inf
    assert b1.fixed[7, 2] == 0.0                                # '- WALL_TEMPERATURE' without a value -> adiabatic
    assert b1.dist.min() > 0 and b0.dist.shape == (b0.kmx + 5, b0.jmx + 5, b0.imx + 5)
    q = helpers.interior(b1.qp, b1)
    assert q[1].min() < 0.5 * f.x_speed_inf                     # the boundary layer came from the restart file
    assert np.all(q[5] == f.tk_inf) and np.all(q[6] == f.tw_inf)   # k, omega are not in the restart list


def test_geometry_invariants(case_mod):
    syn = importlib.import_module("fest3d_b200.synthetic")
    blk = syn.make_duct_blocks(None, n3=(7, 6, 5), turbulence="none", mu_ref=0.0)[0]
    If, Jf, Kf, cells = blk.Ifaces, blk.Jfaces, blk.Kfaces, blk.cells
    for f in (If, Jf, Kf):
        assert np.allclose(np.linalg.norm(f[..., 1:], axis=-1), 1.0, atol=1e-14)
    # closed cells: sum of outward area vectors vanishes
    S = lambda f: f[..., 0:1] * f[..., 1:]
    tot = (S(If)[:, :, 1:] - S(If)[:, :, :-1]) + (S(Jf)[:, 1:, :] - S(Jf)[:, :-1, :]) + (S(Kf)[1:] - S(Kf)[:-1])
    assert np.abs(tot).max() < 1e-16
    # volumes exist on cells 0..imx only, everything else is the placeholder 1.0 (geometry.f90:462-465); the interior
    # cells tile the unit duct (max of two 5-tetrahedra splits: >= the exact volume on warped cells, within 1 %)
    assert np.all(cells[0, :, :, 0] == 1.0) and np.all(cells[..., 0] > 0)
    tot = cells[3:3 + 5, 3:3 + 6, 3:3 + 7, 0].sum()
    assert 1.0 - 1e-12 <= tot < 1.01


def test_mapping_ranges(case_mod):
    assert case_mod._map_range(1, 49) == (1, 48, 1)
    assert case_mod._map_range(49, 1) == (48, 1, -1)


def test_oracle_reproduces_golden_vectors(case_mod, oracle):
    import fixtures
    gv = np.load(os.path.join(GOLDEN, "oracle_vectors.npz"))
    settings = {
        "smoothbump": dict(scheme=dict(time_step_accuracy="RK4"), control=dict(CFL=0.5)),
        "lfp": dict(scheme=dict(scheme_name="slau", interpolant="muscl", time_step_accuracy="RK4"), control=dict(CFL=0.5)),
        "tfp": dict(scheme=dict(scheme_name="ausmUP", interpolant="muscl", time_step_accuracy="RK4"), control=dict(CFL=0.5)),
    }
    for name, st in settings.items():
        blocks = fixtures.load(case_mod, os.path.join(GOLDEN, name), **st)
        w = oracle.OracleWorld(blocks)
        err, res = w.residual(1)
        assert err == 0
        for b in range(len(blocks)):
            assert np.array_equal(res[b], gv["%s_res%d" % (name, b)])
        hist = np.array([w.step(it)[1] for it in range(1, 11)])
        assert np.allclose(hist, gv[name + "_hist"], rtol=1e-13, atol=0)


def test_oracle_block_split_equals_lockstep_order(case_mod, oracle):
    """Two worlds built from the same blocks give identical results (threads per block do not change arithmetic)."""
    syn = importlib.import_module("fest3d_b200.synthetic")
    blocks = syn.make_duct_blocks(None, n3=(8, 6, 5), nb=(2, 1, 1), time_step_accuracy="RK2")
    a, b = oracle.OracleWorld(blocks), oracle.OracleWorld(blocks)
    for it in (1, 2, 3):
        ra, rb = a.step(it)[1], b.step(it)[1]
        assert np.array_equal(ra, rb)
    assert np.array_equal(a.get_state(1), b.get_state(1))


def test_oracle_far_field_turbulence_rule(case_mod, oracle):
    """bc_primitive.f90:700-757: ghost k/omega of a far-field face are decided by the LAST cell of the face loop."""
    syn = importlib.import_module("fest3d_b200.synthetic")
    blk = syn.make_duct_blocks(None, n3=(6, 5, 4), turbulence="sst")[0]
    blk.bc_id = [-8, -8, -6, -6, -6, -6]
    w = oracle.OracleWorld([blk])
    w.residual(1)
    q = w.get_state(0)
    # inflow at imin (u > 0): fixed free-stream k ; outflow at imax: flat copy of the last interior layer
    assert np.all(q[5, 3:3 + 4, 3:3 + 5, 2] == blk.flow.tk_inf)
    assert np.array_equal(q[5, 3:3 + 4, 3:3 + 5, 3 + 6], q[5, 3:3 + 4, 3:3 + 5, 3 + 5])


def test_checkpoint_side_format_roundtrip(tmp_path):
    """The host-side reader / writer of the binary checkpoint format (include/fest3d_gpu.h): 64-byte header, qp with ghosts."""
    import importlib
    ck = importlib.import_module("fest3d_b200.checkpoint")
    rng = np.random.default_rng(7)
    q = rng.standard_normal((7, 4 + 5, 6 + 5, 9 + 5))
    path = str(tmp_path / "a.f3dckpt")
    ck.write_checkpoint(path, q, 17)
    raw = open(path, "rb").read()
    assert len(raw) == 64 + q.size * 8 and raw[:8] == b"F3DCKPT1"
    assert list(np.frombuffer(raw[8:28], dtype="<i4")) == [9, 6, 4, 7, 17]
    hdr, back = ck.read_checkpoint(path)
    assert hdr == dict(imx=9, jmx=6, kmx=4, n_var=7, iter=17) and np.array_equal(back, q)
    open(path, "wb").write(raw[:-8])
    with pytest.raises(ValueError):
        ck.read_checkpoint(path)
    open(path, "wb").write(b"XXXXXXXX" + raw[8:])
    with pytest.raises(ValueError):
        ck.read_checkpoint(path)


def test_bc_file_fixed_values_of_the_second_wave_models(case_mod, tmp_path):
    """read_bc.f90:28-160: '- FIX_tkl v' / '- FIX_tgm v' lines land in their slots of the face they stand under; faces without a line keep the
    free-stream defaults (fill_fixed_values: tkl_inf, tgm_inf)."""
    import importlib
    syn = importlib.import_module("fest3d_b200.synthetic")
    blk = syn.make_duct_blocks(None, n3=(4, 3, 2), turbulence="sst", transition="lctm2015")[0]
    blk.flow.tkl_inf = 3.5e-7
    f = tmp_path / "bc_00.md"
    f.write_text("BOUNDARY CONDITIONS CONFIGURATION\n=================================\n\n# imn\n- FIX_DENSITY 1.3\n- FIX_tgm 0.25\n- FIX_tkl 1e-6\n\n"
                 "# imx\n- COPY_DENSITY\n\n# jmn\n- WALL_TEMPERATURE 300.0\n\n# jmx\n- FIX_tgm 0.75\n\n# kmn\n\n# kmx\n- TOTAL_PRESSURE 123456.0\n\nFIN\n")
    case_mod.read_bc_file(str(f), blk)
    k = case_mod.FIX_KEYS
    assert blk.fixed.shape == (13, 6) and k["FIX_tkl"] == 11 and k["FIX_tgm"] == 12
    assert blk.fixed[k["FIX_DENSITY"], 0] == 1.3 and blk.fixed[k["FIX_DENSITY"], 1] == blk.flow.density_inf
    assert list(blk.fixed[k["FIX_tgm"]]) == [0.25, 1.0, 1.0, 0.75, 1.0, 1.0]           # tgm_inf = 1 (vartypes.f90:258)
    assert blk.fixed[k["FIX_tkl"], 0] == 1e-6 and blk.fixed[k["FIX_tkl"], 3] == 3.5e-7
    assert blk.fixed[k["WALL_TEMPERATURE"], 2] == 300.0 and blk.fixed[k["TOTAL_PRESSURE"], 5] == 123456.0
    # n_var and the initial intermittency (state.f90:260-266, 307-310)
    assert blk.n_var == 8 and case_mod.n_var_of("sa", "lctm2015") == 7 and case_mod.n_var_of("kkl") == 7
    blk.init_state()
    assert blk.qp.shape[0] == 8 and np.all(blk.qp[7] == 1.0)


def test_relative_resnorm_named_norms_and_convergence():
    """resnorm.f90:227-239 (Res_save frozen after iteration 3), :259-360 (named norms), convergence.f90 (tolerance only after iteration 10,
    with its swapped Y / Z momentum entries)."""
    rn = importlib.import_module("fest3d_b200.resnorm")
    h = rn.ResnormHistory("sst")
    base = np.array([1e-3, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0])
    for it in range(1, 13):
        rel = h.update(it, base * 0.5 ** it)
        if it <= 3:
            assert np.all(rel == 1.0)
        else:
            assert np.allclose(rel, 0.5 ** (it - 3), rtol=1e-15)
    a = base * 0.5 ** 12
    assert h.named("Mass_abs") == a[0] and h.named("Continuity_abs") == a[1] and h.named("Omega_abs") == a[7] and h.named("Kl_abs") is None
    assert h.named("Resnorm_abs") == float(np.sqrt((a[1:] ** 2).sum())) and h.named("Viscous_abs") == float(np.sqrt((a[1:6] ** 2).sum()))
    assert h.named("Turbulent_rel") == float(np.sqrt(2.0)) * 0.5 ** 9
    assert h.line(["Mass_abs", "Kl_abs", "TKE_abs"], last_iter=100) == [112, a[0], a[6]]
    assert h.converged(a[1] * 1.01, "Continuity_abs") and not h.converged(a[1] * 0.99, "Continuity_abs")
    assert h.converged(a[3] * 1.01, "Z-mom_abs") and not h.converged(a[3] * 1.01, "Y-mom_abs")          # convergence.f90:31-34 as written
    assert h.converged(1.0, "no such norm") == (h.named("Resnorm_abs") < 1.0)
    early = rn.ResnormHistory("none")
    early.update(10, [0.0] * 6)
    assert not early.converged(1.0, "Resnorm_abs")                                                       # current_iter > 10 only
    # restart: Res_save comes from the restart record and is never overwritten
    r = rn.ResnormHistory("none", previous_res=[1, 2, 2, 2, 2, 2])
    assert np.allclose(r.update(1, [1, 1, 1, 1, 1, 1]), [1, .5, .5, .5, .5, .5])
