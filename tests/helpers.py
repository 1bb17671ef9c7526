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
    geo = importlib.import_module("fest3d_b200.geometry")
    k, j, i = np.meshgrid(np.arange(blk.kmx), np.arange(blk.jmx), np.arange(blk.imx), indexing="ij")
    nodes = np.stack([i * h, j * h, k * h], axis=-1).astype(np.float64)
    blk.nodes = geo.ghost_grid(nodes)
    blk.build_geometry()
    return blk


def interior(q, blk):
    """Interior view of a [nv, kmx+5, jmx+5, imx+5] state."""
    return q[:, 3:3 + blk.kmx - 1, 3:3 + blk.jmx - 1, 3:3 + blk.imx - 1]


def flux_scale(world, b, blk):
    """Per-cell, per-variable flux scale used to normalise residual differences.

    The residual is a difference of six face fluxes, and every upwind face flux is itself a sum of split contributions
    F+ + F- of magnitude  A * rho * (|V| + c) * phi  (phi = 1, velocity scale, total enthalpy, k, omega) that cancel almost
    completely where the face-normal Mach number is small (e.g. the large wall-parallel faces of a boundary-layer cell).
    Round-off (FMA contraction on the GPU vs none in the oracle) is proportional to those split magnitudes, so the scale is
        max( sum_faces |flux_l| ,  sum_faces A_f * rho (|V|+c) phi_l ).
    """
    nv = blk.n_var
    F = world.aux(b, 20, (nv, blk.kmx - 1, blk.jmx - 1, blk.imx))
    G = world.aux(b, 21, (nv, blk.kmx - 1, blk.jmx, blk.imx - 1))
    H = world.aux(b, 22, (nv, blk.kmx, blk.jmx - 1, blk.imx - 1))
    s = (np.abs(F[..., :-1]) + np.abs(F[..., 1:]) + np.abs(G[:, :, :-1, :]) + np.abs(G[:, :, 1:, :])
         + np.abs(H[:, :-1]) + np.abs(H[:, 1:]))
    q = interior(world.get_state(b), blk)
    K, J, I = slice(3, 3 + blk.kmx - 1), slice(3, 3 + blk.jmx - 1), slice(3, 3 + blk.imx - 1)
    K1, J1, I1 = slice(4, 4 + blk.kmx - 1), slice(4, 4 + blk.jmx - 1), slice(4, 4 + blk.imx - 1)
    area = (blk.Ifaces[K, J, I, 0] + blk.Ifaces[K, J, I1, 0] + blk.Jfaces[K, J, I, 0] + blk.Jfaces[K, J1, I, 0]
            + blk.Kfaces[K, J, I, 0] + blk.Kfaces[K1, J, I, 0])
    gm = blk.flow.gm
    rho, p = q[0], q[4]
    c = np.sqrt(gm * p / rho)
    vm = np.sqrt(q[1] ** 2 + q[2] ** 2 + q[3] ** 2)
    Ht = gm / (gm - 1.0) * p / rho + 0.5 * vm ** 2
    m = area * rho * (vm + c)
    phi = [np.ones_like(m), vm + c, vm + c, vm + c, Ht] + [np.abs(q[v]) for v in range(5, nv)]
    ac = np.stack([m * f for f in phi])
    return np.maximum(s, ac)


def residual_parity(r_gpu, r_orc, scale):
    """max over cells of |dR| / flux scale (per variable).  Variables whose flux scale is identically zero must agree exactly."""
    out = []
    for v in range(r_gpu.shape[0]):
        sc = scale[v]
        floor = max(sc.max(), 1e-300) * 1e-6   # cells with (near-)vanishing flux take the block's scale
        out.append(float((np.abs(r_gpu[v] - r_orc[v]) / np.maximum(sc, floor)).max()))
    return out


def state_rel_diff(a, b):
    out = []
    for v in range(a.shape[0]):
        den = max(np.abs(b[v]).max(), 1e-300)
        out.append(float(np.abs(a[v] - b[v]).max() / den))
    return out


def boundary_mass_flux_scale(world, blocks):
    """sum over all blocks and all six block faces of |mass flux| of the fluxes the oracle holds (F, G, H of the last stage):
    the scale the round-off of Res_abs(0) = |sum of signed boundary mass fluxes| (resnorm.f90:190-198) lives on."""
    tot = 0.0
    for b, blk in enumerate(blocks):
        nv = blk.n_var
        F = world.aux(b, 20, (nv, blk.kmx - 1, blk.jmx - 1, blk.imx))
        G = world.aux(b, 21, (nv, blk.kmx - 1, blk.jmx, blk.imx - 1))
        H = world.aux(b, 22, (nv, blk.kmx, blk.jmx - 1, blk.imx - 1))
        tot += (np.abs(F[0, :, :, 0]).sum() + np.abs(F[0, :, :, -1]).sum() + np.abs(G[0, :, 0, :]).sum() + np.abs(G[0, :, -1, :]).sum()
                + np.abs(H[0, 0]).sum() + np.abs(H[0, -1]).sum())
    return float(tot)
