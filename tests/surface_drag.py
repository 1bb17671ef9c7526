"""Restatement of the reference's flat-plate post-processing (tests/Lfp/pp/Surface.py:87-153, tests/Tfp/pp/Surface.py: the same with
another reference area): drag coefficient of the no-slip walls (BC id -5) from the cell-centred velocity gradients, viscosity and
pressure of the first cell row and the wall face's normal and area.

    tau_ij = mu ((du_i/dx_j + du_j/dx_i) - 2/3 div u delta_ij);  F = tau . n
    C_d = [ sum_faces (F . n_drag) A / q_dyn  +  sum_faces (-c_p n . n_drag A) ] / A_ref,   q_dyn = 1/2 rho_inf |V_inf|^2

Like the reference script it handles walls on j faces (its face data are the j normals whatever the face): both flat-plate cases
have their wall at jmin of block 1.  Inputs are what the reference writes to its result files (Dudx .. Dwdz, Mu, Pressure at cell
centres); here they come from the device through fest3d_gpu_get_aux / get_state."""
import numpy as np

A_REF = {"lfp": 0.05, "tfp": 2 * 0.04}          # Surface.py:144 of each case
CD_REPORT = {"lfp": 1.329e-3, "tfp": 2.872e-3}  # tests/Report.txt:20-23, 33-36 (the reference's own run)
CD_EXPECTED = {"lfp": 1.33e-3, "tfp": 2.90e-3}  # the scripts' target, tolerance 1 % / 2 %


def wall_drag(blocks, grads, mus, states, a_ref, n_drag=(1.0, 0.0, 0.0)):
    """blocks: BlockSetup list; grads[b] = (gx, gy, gz) each [n_grad, kmx+1, jmx+1, imx+1] (cells 0..imx); mus[b] = mu
    [kmx+5, jmx+5, imx+5]; states[b] = qp [nv, kmx+5, jmx+5, imx+5].  Returns C_d."""
    f = blocks[0].flow
    qdyn = 0.5 * f.density_inf * (f.x_speed_inf ** 2 + f.y_speed_inf ** 2 + f.z_speed_inf ** 2)
    cd = 0.0
    for b, blk in enumerate(blocks):
        for face in (2, 3):                      # jmin, jmax
            if blk.bc_id[face] != -5:
                continue
            j = 1 if face == 2 else blk.jmx - 1  # first / last interior cell row
            jf = 1 if face == 2 else blk.jmx     # the wall face
            sgn = 1.0 if face == 2 else -1.0
            K, I = slice(3, 3 + blk.kmx - 1), slice(3, 3 + blk.imx - 1)
            A = blk.Jfaces[K, jf + 2, I, 0]
            n = sgn * blk.Jfaces[K, jf + 2, I, 1:4]
            gx, gy, gz = grads[b]
            Kc, Ic = slice(1, blk.kmx), slice(1, blk.imx)        # gradient arrays start at cell 0
            dudx, dudy, dudz = gx[0, Kc, j, Ic], gy[0, Kc, j, Ic], gz[0, Kc, j, Ic]
            dvdx, dvdy, dvdz = gx[1, Kc, j, Ic], gy[1, Kc, j, Ic], gz[1, Kc, j, Ic]
            dwdx, dwdy, dwdz = gx[2, Kc, j, Ic], gy[2, Kc, j, Ic], gz[2, Kc, j, Ic]
            mu = mus[b][K, j + 2, I]
            p = states[b][4, K, j + 2, I]
            delv = dudx + dvdy + dwdz
            txx = mu * ((dudx + dudx) - 2.0 * delv / 3.0)
            tyy = mu * ((dvdy + dvdy) - 2.0 * delv / 3.0)
            tzz = mu * ((dwdz + dwdz) - 2.0 * delv / 3.0)
            txy = mu * (dudy + dvdx); tyz = mu * (dwdy + dvdz); txz = mu * (dudz + dwdx)
            nx, ny, nz = n[..., 0], n[..., 1], n[..., 2]
            Fx = txx * nx + txy * ny + txz * nz
            Fy = txy * nx + tyy * ny + tyz * nz
            Fz = txz * nx + tyz * ny + tzz * nz
            cp = (p - f.pressure_inf) / qdyn
            cd += np.sum((Fx * n_drag[0] + Fy * n_drag[1] + Fz * n_drag[2]) * A / qdyn)
            cd += np.sum(-cp * nx * A * n_drag[0] + -cp * ny * A * n_drag[1] + -cp * ny * A * n_drag[2])   # (sic: Surface.py:132)
    return float(cd / a_ref)


def device_wall_drag(solver, blocks, case):
    """C_d of the state the device holds (gradients and mu through the fest3d_gpu_get_aux views)."""
    grads, mus, states = [], [], []
    for gb, blk in zip(solver.blocks, blocks):
        ng = 6 if blk.n_var == 7 else (5 if blk.n_var == 6 else 4)
        shp = (ng, blk.kmx + 1, blk.jmx + 1, blk.imx + 1)
        grads.append(tuple(gb.aux(30 + d, shp) for d in range(3)))
        mus.append(gb.aux(1, (blk.kmx + 5, blk.jmx + 5, blk.imx + 5)))
        states.append(gb.get_state())
    return wall_drag(blocks, grads, mus, states, A_REF[case])
