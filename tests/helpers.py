"""Shared helpers for the tests (block builders, comparators)."""
import importlib

import numpy as np


def blank_block(case_mod, imx, jmx, kmx, bc_id=None, interpolant="muscl", scheme_name="ausm", turbulence="none",
                mu_ref=0.0, mu_variation="constant", time_step_accuracy="none", **flow_kw):
    sch = case_mod.Scheme(scheme_name=scheme_name, interpolant=interpolant, limiter=(0, 0, 0), tlimiter=(1, 1, 1),
                          turbulence=turbulence, time_step_accuracy=time_step_accuracy)
    fl = case_mod.Flow(mu_ref=mu_ref, mu_variation=mu_variation, **flow_kw)
    fl.derive(turbulence)
    blk = case_mod.BlockSetup(imx=imx, jmx=jmx, kmx=kmx, bc_id=list(bc_id or [3] * 6), scheme=sch, flow=fl,
                              control=case_mod.Control(CFL=1.0))
    blk.default_maps()
    blk.fill_fixed_defaults()
    blk.cells = np.zeros((kmx + 5, jmx + 5, imx + 5, 4)); blk.cells[..., 0] = 1.0
    blk.Ifaces = np.zeros((kmx + 5, jmx + 5, imx + 6, 4))
    blk.Jfaces = np.zeros((kmx + 5, jmx + 6, imx + 5, 4))
    blk.Kfaces = np.zeros((kmx + 6, jmx + 5, imx + 5, 4))
    if turbulence != "none":
        blk.dist = np.ones((kmx + 5, jmx + 5, imx + 5))
    blk.init_state()
    return blk


def unit_cube_geometry(blk, h=1.0):
    geo = importlib.import_module("fest-3d_b200.geometry")
    k, j, i = np.meshgrid(np.arange(blk.kmx), np.arange(blk.jmx), np.arange(blk.imx), indexing="ij")
    nodes = np.stack([i * h, j * h, k * h], axis=-1).astype(np.float64)
    blk.nodes = geo.ghost_grid(nodes)
    blk.build_geometry()
    return blk


def interior(q, blk):
    """Interior view of a [nv, kmx+5, jmx+5, imx+5] state."""
    return q[:, 3:3 + blk.kmx - 1, 3:3 + blk.jmx - 1, 3:3 + blk.imx - 1]
