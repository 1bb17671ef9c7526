"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
PARITY UNPINNED against a runnable reference (see oracle_abi.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NFIX = 13


class OracleConfig(C.Structure):
    _fields_ = [
        ("imx", C.c_int), ("jmx", C.c_int), ("kmx", C.c_int), ("n_var", C.c_int),
        ("scheme", C.c_int), ("interpolant", C.c_int), ("turbulence", C.c_int), ("transition", C.c_int),
        ("time_accuracy", C.c_int), ("time_stepping", C.c_int),
        ("limiter", C.c_int * 3), ("tlimiter", C.c_int * 3), ("pb_switch", C.c_int * 3),
        ("accur", C.c_int), ("mu_variation", C.c_int),
        ("bc_id", C.c_int * 6), ("pbc_id", C.c_int * 6), ("dir_switch", C.c_int * 6), ("otherface", C.c_int * 6),
        ("plo", (C.c_int * 2) * 6), ("phi", (C.c_int * 2) * 6), ("pdir", (C.c_int * 2) * 6),
        ("block_id", C.c_int), ("n_blocks", C.c_int),
        ("CFL", C.c_double), ("global_time_step", C.c_double),
        ("gm", C.c_double), ("R_gas", C.c_double), ("mu_ref", C.c_double), ("T_ref", C.c_double),
        ("Sutherland_temp", C.c_double), ("Pr", C.c_double), ("tPr", C.c_double),
        ("density_inf", C.c_double), ("x_speed_inf", C.c_double), ("y_speed_inf", C.c_double),
        ("z_speed_inf", C.c_double), ("pressure_inf", C.c_double),
        ("tk_inf", C.c_double), ("tw_inf", C.c_double), ("vel_mag", C.c_double), ("MInf", C.c_double), ("tv_inf", C.c_double), ("tu_inf", C.c_double), ("tkl_inf", C.c_double), ("tgm_inf", C.c_double),
        ("fixed", (C.c_double * 6) * NFIX),
    ]


def fill_config(cfg, blk, enums):
    """Fill a ctypes config (oracle or GPU -- both use the same field names) from a BlockSetup."""
    SCHEMES, INTERPOLANTS, TURBULENCE, TRANSITION, TIME_ACCURACY = enums
    s, f, c = blk.scheme, blk.flow, blk.control
    cfg.imx, cfg.jmx, cfg.kmx, cfg.n_var = blk.imx, blk.jmx, blk.kmx, blk.n_var
    cfg.scheme = SCHEMES[s.scheme_name]; cfg.interpolant = INTERPOLANTS[s.interpolant]
    cfg.turbulence = TURBULENCE[s.turbulence]; cfg.transition = TRANSITION[s.transition]
    cfg.time_accuracy = TIME_ACCURACY[s.time_step_accuracy]
    cfg.time_stepping = 1 if s.time_stepping_method == "g" else 0
    for d in range(3):
        cfg.limiter[d] = s.limiter[d]; cfg.tlimiter[d] = s.tlimiter[d]; cfg.pb_switch[d] = s.pb_switch[d]
    cfg.accur = s.accur
    cfg.mu_variation = 1 if f.mu_variation == "sutherland_law" else 0
    for i in range(6):
        cfg.bc_id[i] = blk.bc_id[i]; cfg.pbc_id[i] = blk.pbc_id[i]
        cfg.dir_switch[i] = blk.dir_switch[i]; cfg.otherface[i] = blk.otherface[i]
        for t in range(2):
            cfg.plo[i][t] = blk.plo[i][t]; cfg.phi[i][t] = blk.phi[i][t]; cfg.pdir[i][t] = blk.pdir[i][t]
    cfg.block_id, cfg.n_blocks = blk.block_id, blk.n_blocks
    cfg.CFL = c.CFL; cfg.global_time_step = s.global_time_step
    for k in ("gm", "R_gas", "mu_ref", "T_ref", "Sutherland_temp", "Pr", "tPr", "density_inf", "x_speed_inf",
              "y_speed_inf", "z_speed_inf", "pressure_inf", "tk_inf", "tw_inf", "vel_mag", "MInf", "tv_inf", "tu_inf", "tkl_inf", "tgm_inf"):
        setattr(cfg, k, getattr(f, k))
    for sl in range(NFIX):
        for i in range(6):
            cfg.fixed[sl][i] = float(blk.fixed[sl, i])
    return cfg


_libs = {}


def build(fast=False):
    name = "liboracle_fast.so" if fast else "liboracle.so"
    path = os.path.join(HERE, name)
    srcs = [os.path.join(HERE, f) for f in ("oracle_inviscid.cpp", "oracle_viscous.cpp", "oracle_api.cpp", "oracle_core.hpp", "oracle_abi.h")]
    if (not os.path.exists(path)) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, name], stdout=subprocess.DEVNULL)
    return path


def lib(fast=False):
    if fast in _libs:
        return _libs[fast]
    L = C.CDLL(build(fast))
    dp = C.POINTER(C.c_double)
    L.oracle_create.restype = C.c_void_p
    L.oracle_create.argtypes = [C.c_int, C.POINTER(OracleConfig)]
    L.oracle_destroy.argtypes = [C.c_void_p]
    L.oracle_set_geometry.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, dp, dp]
    L.oracle_set_state.argtypes = [C.c_void_p, C.c_int, dp]
    L.oracle_get_state.argtypes = [C.c_void_p, C.c_int, dp]
    L.oracle_residual.argtypes = [C.c_void_p, C.c_int]
    L.oracle_get_residue.argtypes = [C.c_void_p, C.c_int, dp]
    L.oracle_step.argtypes = [C.c_void_p, C.c_int, dp]
    L.oracle_get_aux.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
    L.oracle_kat_flux.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, dp, dp, dp, C.c_int, dp]
    L.oracle_kat_states.argtypes = [C.c_int, C.c_int, dp, dp, C.c_int, dp, dp]
    L.oracle_kat_residue.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp]
    L.oracle_kat_residue.restype = None
    L.oracle_find_wall_dist.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, C.c_longlong, dp]
    L.oracle_find_wall_dist.restype = None
    _libs[fast] = L
    return L


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class OracleWorld:
    """All blocks of a case stepped in lock step (one 'MPI rank' per block)."""

    def __init__(self, blocks, fast=False):
        from importlib import import_module
        case = import_module("fest3d_b200.case")
        enums = (case.SCHEMES, case.INTERPOLANTS, case.TURBULENCE, case.TRANSITION, case.TIME_ACCURACY)
        self.L = lib(fast)
        self.blocks = blocks
        cfgs = (OracleConfig * len(blocks))()
        for i, b in enumerate(blocks):
            fill_config(cfgs[i], b, enums)
        self.h = self.L.oracle_create(len(blocks), cfgs)
        for i, b in enumerate(blocks):
            dist = np.ascontiguousarray(b.dist) if b.dist is not None else None
            self.L.oracle_set_geometry(self.h, i, _p(b.cells), _p(b.Ifaces), _p(b.Jfaces), _p(b.Kfaces), _p(dist))
            self.set_state(i, b.qp)

    def __del__(self):
        try:
            self.L.oracle_destroy(self.h)
        except Exception:
            pass

    def set_state(self, b, qp):
        q = np.ascontiguousarray(qp, dtype=np.float64)
        self.L.oracle_set_state(self.h, b, _p(q))

    def get_state(self, b):
        blk = self.blocks[b]
        q = np.empty((blk.n_var, blk.kmx + 5, blk.jmx + 5, blk.imx + 5))
        self.L.oracle_get_state(self.h, b, _p(q))
        return q

    def residual(self, current_iter=1):
        err = self.L.oracle_residual(self.h, current_iter)
        out = []
        for i, blk in enumerate(self.blocks):
            r = np.empty((blk.n_var, blk.kmx - 1, blk.jmx - 1, blk.imx - 1))
            self.L.oracle_get_residue(self.h, i, _p(r))
            out.append(r)
        return err, out

    def step(self, current_iter):
        nv = self.blocks[0].n_var
        res = np.zeros(nv + 1)
        err = self.L.oracle_step(self.h, current_iter, _p(res))
        return err, res

    def aux(self, b, which, shape):
        a = np.empty(shape)
        rc = self.L.oracle_get_aux(self.h, b, which, _p(a))
        assert rc == 0
        return a
