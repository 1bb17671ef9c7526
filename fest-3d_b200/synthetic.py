"""Synthetic structured duct grids for the throughput configurations (BASELINE.json configs 4 and 5).

A unit-cube duct [0,1]^3 is cut into nbx x nby x nbz blocks of n^3 cells each.  Nodes sit on a uniform lattice plus
a smooth deterministic warp  delta = 0.1 h sin(2 pi x) sin(2 pi y) sin(2 pi z)  (zero on the duct walls and on every
block interface plane of a 2x2x2 split), so metrics are non-trivial.  Outer faces: imin subsonic inlet (-3), imax
subsonic outlet (-4), the four side faces no-slip walls (-5); inner faces are interfaces (id = neighbour block,
dir_switch 0).  Flow = tests/Tfp/system/flow.md of the reference.  The wall distance is the analytic minimum distance
to the four walls at the cell centres.  The state is free stream times (1 + 1e-3 * closed-form trigonometric
perturbation): no RNG, nothing constant-foldable, no 0/0 in the limiters.
"""
from __future__ import annotations

import copy

import numpy as np

from . import case as case_mod
from . import geometry as geo

TFP_FLOW = dict(density_inf=1.17659, x_speed_inf=69.445, y_speed_inf=0.0, z_speed_inf=0.0, pressure_inf=101325.0,
                tu_inf=0.03873, mu_ratio_inf=0.01, tgm_inf=1.0, mu_ref=1.63416585e-05, mu_variation="sutherland_law",
                T_ref=300.0, Sutherland_temp=110.5, Pr=0.72, tPr=0.9, gm=1.4, R_gas=287.0)


def make_duct_blocks(n, nb=(1, 1, 1), scheme_name="ausm", interpolant="muscl", turbulence="sst", time_step_accuracy="none",
                     CFL=0.5, limiter=(1, 1, 1), tlimiter=(1, 1, 1), only_blocks=None, mu_ref=None, n3=None, transition="none"):
    """Return the list of BlockSetup for the duct.  ``n`` = cells per block edge (or n3 = (ni,nj,nk)).
    ``only_blocks``: build just these block ids (a rank builds only what it owns)."""
    ni, nj, nk = n3 if n3 is not None else (n, n, n)
    nbx, nby, nbz = nb
    n_blocks = nbx * nby * nbz
    flow_kw = dict(TFP_FLOW)
    if turbulence == "none" and mu_ref is None:
        pass
    if mu_ref is not None:
        flow_kw["mu_ref"] = mu_ref
    sch = case_mod.Scheme(scheme_name=scheme_name, interpolant=interpolant, limiter=tuple(limiter), tlimiter=tuple(tlimiter),
                          turbulence=turbulence, transition=transition, time_step_accuracy=time_step_accuracy, time_stepping_method="l", accur=0)
    fl = case_mod.Flow(**flow_kw).derive(turbulence)
    ctl = case_mod.Control(CFL=CFL)
    Nx, Ny, Nz = nbx * ni, nby * nj, nbz * nk     # global cell counts
    hx, hy, hz = 1.0 / Nx, 1.0 / Ny, 1.0 / Nz
    h = min(hx, hy, hz)
    blocks = []
    for bz in range(nbz):
        for by in range(nby):
            for bx in range(nbx):
                bid = bx + nbx * (by + nby * bz)
                if only_blocks is not None and bid not in only_blocks:
                    continue

                def nid(dx, dy, dz):
                    x, y, z = bx + dx, by + dy, bz + dz
                    return x + nbx * (y + nby * z)
                ids = [-3 if bx == 0 else nid(-1, 0, 0), -4 if bx == nbx - 1 else nid(1, 0, 0),
                       -5 if by == 0 else nid(0, -1, 0), -5 if by == nby - 1 else nid(0, 1, 0),
                       -5 if bz == 0 else nid(0, 0, -1), -5 if bz == nbz - 1 else nid(0, 0, 1)]
                blk = case_mod.BlockSetup(imx=ni + 1, jmx=nj + 1, kmx=nk + 1, bc_id=ids, scheme=copy.copy(sch), flow=copy.copy(fl),
                                          control=copy.copy(ctl), block_id=bid, n_blocks=n_blocks)
                blk.default_maps()
                blk.fill_fixed_defaults()
                # nodes of this block (interior node lattice), then the reference's ghost extrapolation
                gi = (bx * ni + np.arange(ni + 1)) * hx
                gj = (by * nj + np.arange(nj + 1)) * hy
                gk = (bz * nk + np.arange(nk + 1)) * hz
                Z, Y, X = np.meshgrid(gk, gj, gi, indexing="ij")
                w = 0.1 * h * np.sin(2 * np.pi * X) * np.sin(2 * np.pi * Y) * np.sin(2 * np.pi * Z)
                nodes = np.stack([X + w, Y + 0.5 * w, Z - 0.7 * w], axis=-1)
                blk.nodes = geo.ghost_grid(nodes)
                blk.build_geometry()
                cx, cy, cz = blk.cells[..., 1], blk.cells[..., 2], blk.cells[..., 3]
                if turbulence != "none":
                    d = np.minimum(np.minimum(np.abs(cy), np.abs(1.0 - cy)), np.minimum(np.abs(cz), np.abs(1.0 - cz)))
                    blk.dist = np.ascontiguousarray(np.maximum(d, 0.25 * h))
                blk.init_state()
                # closed-form perturbation of the free stream (ghost cells included; they are refilled anyway)
                s1 = np.sin(2 * np.pi * cx) * np.cos(2 * np.pi * cy) * np.cos(2 * np.pi * cz)
                s2 = np.cos(4 * np.pi * cx + 0.3) * np.sin(2 * np.pi * cy + 0.1) * np.cos(2 * np.pi * cz - 0.2)
                s3 = np.sin(2 * np.pi * cx - 0.5) * np.sin(4 * np.pi * cy) * np.sin(2 * np.pi * cz + 0.4)
                q = blk.qp
                q[0] *= 1 + 1e-3 * s1
                q[1] *= 1 + 1e-3 * s2
                q[2] = fl.x_speed_inf * 1e-3 * s3
                q[3] = fl.x_speed_inf * 1e-3 * s1 * s2
                q[4] *= 1 + 1e-3 * s3
                if blk.n_var >= 7:
                    q[5] *= 1 + 1e-3 * s2
                    q[6] *= 1 + 1e-3 * s1
                    if blk.n_var == 8:      # intermittency in (0, 1): the explicit update never rewrites it (update.f90:349-362)
                        q[7] = 0.55 + 0.4 * s3
                elif blk.n_var == 6:
                    q[5] *= 1 + 1e-3 * s2
                blocks.append(blk)
    return blocks
