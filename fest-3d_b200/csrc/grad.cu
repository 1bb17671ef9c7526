// Gradient / viscosity kernels of the STAGED form of the viscous path (the default, F3D_GRADIENTS=staged): Green-Gauss cell gradients of
// (u,v,w,T[,k,omega | nu-tilde][,gamma]) on cells 0..imx x 0..jmx x 0..kmx with the molecular (Sutherland) and eddy viscosity / SST
// blending function F1 of the same cell, then the ghost-gradient rule and the ghost mu_t / F1 copies on physical faces.  The sweep
// (sweep3_kernel.cuh) stages these arrays with TMA.  With F3D_GRADIENTS=fused the sweep computes the same quantities in shared memory
// (fused_kernel.cuh) and these kernels only serve the views of fest3d_gpu_get_aux (which = 1..3, 30..32).
//
// Reference: src/gradients.f90:276-402 (evaluate_all_gradients), :405-482 (compute_gradient_G), :486-676
// (apply_gradient_bc, incl. the Ifaces-shaped dummy that mis-indexes Jfaces/Kfaces -- those records are gathered on the
// host into ctx->gbc with the reference's linear offset); src/viscosity.f90:109-138 (Sutherland), :343-388 (sst),
// :215-263 (sst2003), :265-279 / :390-404 (lctm2015 F1), :408-465 (ghost mu_t / F1), :469-533 (kkl); src/CC.f90:73-200 (lctm2015 fields).
#include "ctx.hpp"
#include "physics.cuh"
#include <algorithm>
#include <cstdlib>

namespace f3d {

#ifdef F3D_GRAD_UNALIGNED
constexpr int G_ALIGN = 0;
#else
constexpr int G_ALIGN = 15;
#endif
// Green-Gauss gradient of (u, v, w, T [, k, omega | nu-tilde] [, gamma]) at cell c (gradients.f90:405-482)
template <int NG>
__device__ __forceinline__ void green_gauss(const Params& P, const double* __restrict__ q, const double* __restrict__ temp, const double* __restrict__ geom,
                                            long long c, double (&g)[NG][3], double& nan_probe) {
  const Layout& L = P.L;
  const long long fs = L.fs, sj = L.sj, sk = L.sk;
  const double* gI = geom + (long long)G_IA * fs;
  const double* gJ = geom + (long long)G_JA * fs;
  const double* gK = geom + (long long)G_KA * fs;
  const double AIl = gI[c], AIh = gI[c + 1], AJl = gJ[c], AJh = gJ[c + sj], AKl = gK[c], AKh = gK[c + sk];
  const double ivol2 = rcp64(2 * geom[(long long)G_VOL * fs + c]);
  const bool zgrad = L.kmx > 2;   // gradqp_z = 0 when kmx == 2 (gradients.f90:328-336)
  // face weights n*A once per face and direction (18 products), then 6 FMAs per gradient component: branch-free so the
  // NG*3 independent chains interleave
  double wl[3][3], wh[3][3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    wl[0][d] = gI[(1 + d) * fs + c] * AIl; wl[1][d] = gJ[(1 + d) * fs + c] * AJl; wl[2][d] = gK[(1 + d) * fs + c] * AKl;
    wh[0][d] = gI[(1 + d) * fs + c + 1] * AIh; wh[1][d] = gJ[(1 + d) * fs + c + sj] * AJh; wh[2][d] = gK[(1 + d) * fs + c + sk] * AKh;
  }
#pragma unroll
  for (int cc = 0; cc < NG; ++cc) {
    const double* __restrict__ var = (cc == 3) ? temp : (q + (long long)(cc + 1) * fs);
    const double v0 = var[c];
    const double sIl = var[c - 1] + v0, sJl = var[c - sj] + v0, sKl = var[c - sk] + v0;
    const double sIh = var[c + 1] + v0, sJh = var[c + sj] + v0, sKh = var[c + sk] + v0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      double r = (-sIl * wl[0][d] - sJl * wl[1][d] - sKl * wl[2][d] + sIh * wh[0][d] + sJh * wh[1][d] + sKh * wh[2][d]) * ivol2;
      if (d == 2) r = zgrad ? r : 0.0;
      nan_probe += r;
      g[cc][d] = r;
    }
  }
}

// molecular viscosity (mu_ref when constant, viscosity.f90:527; Sutherland :109-138)
__device__ __forceinline__ double molecular_viscosity(const Params& P, const double* __restrict__ q, long long c, int* err) {
  double mu = P.mu_ref;
  if (P.mu_variation == 1) {
    const double T = q[4 * P.L.fs + c] * rcp64(q[c] * P.R_gas);
    const double tr = T / P.T_ref;   // (T/T_ref)**1.5 = tr*sqrt(tr): <= 1 ulp from pow, 8x cheaper
    mu = P.mu_ref * (tr * sqrt(tr)) * ((P.T_ref + P.Sutherland_temp) * rcp64(T + P.Sutherland_temp));
    if (isnan(mu)) atomicOr(err, F3D_ERR_NAN_VISCOSITY);
  }
  return mu;
}

// eddy viscosity and blending function of cell c from its gradients g and molecular viscosity mu
// (viscosity.f90:149-163 sa, :215-263 sst2003, :343-388 sst, :265-279 / :390-404 lctm2015, :469-484 kkl)
template <int NG>
__device__ __forceinline__ void eddy_viscosity(const Params& P, const double* __restrict__ q, const double* __restrict__ geom, long long c,
                                               const double (&g)[NG][3], double mu, double& mut, double& F1) {
  const Layout& L = P.L;
  const long long fs = L.fs;
  constexpr bool TWO_EQ = (NG >= 6);   // sst, sst2003, kkl (6); sst / sst2003 with the intermittency of lctm2015 (7)
  mut = 0.0; F1 = 0.0;
  if (NG == 5) {   // Spalart-Allmaras: mu_t = rho*tv*fv1
    const double tv = q[5 * fs + c], density = q[c];
    const double xi = tv * density / mu;
    const double fv1 = (pow3(xi)) / ((pow3(xi)) + (pow3(kCv1)));
    mut = density * tv * fv1;
  }
  if (TWO_EQ && P.kkl) {   // k-kL: mu_t = cmu^(1/4) rho kL / max(sqrt(k), 1e-20), 0 below 1e-14; no F1
    const double density = q[c], tk = q[5 * fs + c], tkl = q[6 * fs + c];
    double m = kKklCmu25 * density * tkl / (fmax(sqrt(tk), 1.e-20));
    if (tkl < 1.e-14 || tk < 1.e-14) m = 0.0;
    mut = m;
  } else if (TWO_EQ) {
    const double density = q[c], tk = q[5 * fs + c], tw = q[6 * fs + c];
    const double d = geom[(long long)G_DIST * fs + c];
    const double var1 = sqrt(tk) * rcp64(kBstar * tw * d);
    const double var2 = 500 * (mu * rcp64(density)) * rcp64((d * d) * tw);
    const double arg2 = dmax(2 * var1, var2);
    const double Fb = tanh(arg2 * arg2);
    double rate;
    if (P.turbulence == F3D_TURB_SST) {
      const double wx = g[2][1] - g[1][2], wy = g[0][2] - g[2][0], wz = g[1][0] - g[0][1];
      rate = sqrt(wx * wx + wy * wy + wz * wz);
    } else {
      const double sxx = g[0][0], syy = g[1][1], szz = g[2][2];
      const double syz = g[2][1] + g[1][2], szx = g[0][2] + g[2][0], sxy = g[1][0] + g[0][1];
      rate = sqrt((2.0 * (sxx * sxx)) + (2.0 * (syy * syy)) + (2.0 * (szz * szz)) + syz * syz + szx * szx + sxy * sxy);
    }
    const double NUM = density * kA1 * tk;
    const double DENOM = dmax(dmax((kA1 * tw), rate * Fb), P.mut_floor);
    mut = NUM * rcp64(DENOM);
    const double CD = dmax(2 * density * kSigmaW2 * (g[4][0] * g[5][0] + g[4][1] * g[5][1] + g[4][2] * g[5][2]) * rcp64(tw), P.mut_floor);
    const double right = 4 * (density * kSigmaW2 * tk) * rcp64(CD * (d * d));
    const double left = dmax(var1, var2);
    const double arg1 = dmin(left, right);
    F1 = tanh((arg1 * arg1) * (arg1 * arg1));
    if (NG == 7) {
      // viscosity.f90:265-279 / :390-404 "modified blending function (Menter 2015)".  KEPT DEFECT: its loop reuses the scalars `density`
      // and `tk` the loop above left behind -- those of its last cell (imx, jmx, kmx), a corner ghost cell -- for every cell.
      const long long cl = L.idx(L.imx, L.jmx, L.kmx);
      const double x = (q[cl] * d * sqrt(q[5 * fs + cl]) / mu) / 120;
      const double x2 = x * x, x4 = x2 * x2;
      F1 = fmax(F1, exp(-(x4 * x4)));
    }
  }
}

// add_sst_source_lctm2015 (source.f90:273-463) of one interior cell: S_k V, S_omega V, S_gamma V.  dvdy is the CC.f90 field (k_dvdy).
__device__ __forceinline__ void lctm_sources(const Params& P, const double (&g)[7][3], double density, double tk, double tw, double gm_, double mu_c,
                                             double mut, double F1c, double dvdy, double dist_c, double volc, double& Sk, double& Sw, double& Sg) {
  const double wx = g[2][1] - g[1][2], wy = g[0][2] - g[2][0], wz = g[1][0] - g[0][1];
  const double vort = sqrt(wx * wx + wy * wy + wz * wz);
  const double syz = g[2][1] + g[1][2], szx = g[0][2] + g[2][0], sxy = g[1][0] + g[0][1];
  const double strain = sqrt(((syz * syz) + (szx * szx) + (sxy * sxy) + 2 * (g[0][0] * g[0][0]) + 2 * (g[1][1] * g[1][1]) + 2 * (g[2][2] * g[2][2])));
  double CD = 2 * density * kSigmaW2 * (g[4][0] * g[5][0] + g[4][1] * g[5][1] + g[4][2] * g[5][2]) / tw;
  CD = fmax(CD, P.cd_floor);
  const double gama = P.gama1 * F1c + P.gama2 * (1. - F1c);
  const double beta = kBeta1 * F1c + kBeta2 * (1. - F1c);
  const double D_k = kBstar * density * tw * tk;
  const double D_w = beta * density * (tw * tw);
  const double divergence = g[0][0] + g[1][1] + g[2][2];
  double P_k = mut * (vort * strain) - ((2.0 / 3.0) * density * tk * divergence);
  P_k = fmin(P_k, P.pk_limiter * D_k);
  const double P_w = (density * gama / mut) * P_k;
  const double lamda = (1. - F1c) * CD;
  double lamd = (-7.57e-3) * (dvdy * dist_c * dist_c * density / mu_c) + 0.0128;
  lamd = fmin(fmax(lamd, -1.0), 1.0);
  double Fpg = (lamd >= 0.0) ? fmin(1.0 + 14.68 * lamd, 1.5) : fmin(1.0 - 7.34 * lamd, 3.0);
  Fpg = fmax(Fpg, 0.0);
  const double TuL = fmin(100.0 * sqrt(2.0 * tk / 3.0) / (tw * dist_c), 100.0);
  const double Re_theta = 100.0 + 1000.0 * exp(-TuL * Fpg);
  const double Rev = density * dist_c * dist_c * strain / mu_c;
  const double RT = density * tk / (mu_c * tw);
  const double hr = 0.5 * RT;
  const double Fturb = exp(-((hr * hr) * (hr * hr)));
  const double Fonset1 = Rev / (2.2 * Re_theta);
  const double Fonset2 = fmin(Fonset1, 2.0);
  const double r35 = RT / 3.5;
  const double Fonset3 = fmax(1.0 - (r35 * r35 * r35), 0.0);
  const double Fonset = fmax(Fonset2 - Fonset3, 0.0);
  const double P_gm = 100 * density * strain * gm_ * (1.0 - gm_) * Fonset;
  const double D_gm = 0.06 * density * vort * gm_ * Fturb * ((50.0 * gm_) - 1.0);
  const double Fon_lim = fmin(fmax((Rev / (2.2 * 1100.0)) - 1.0, 0.0), 3.0);
  const double Pk_lim = 5 * fmax(gm_ - 0.2, 0.0) * (1.0 - gm_) * Fon_lim * fmax(3 * mu_c - mut, 0.0) * strain * vort;
  Sk = (gm_ * P_k - fmax(gm_, 0.1) * D_k + Pk_lim) * volc;
  Sw = (P_w - D_w + lamda) * volc;
  Sg = (P_gm - D_gm) * volc;
}

// One thread per cell of 0..imx x 0..jmx x 0..kmx: gradient, molecular and eddy viscosity, F1 of the cell itself
template <int NG>
__global__ void __launch_bounds__(128, 5) k_gradients(const Params P, const double* __restrict__ q, const double* __restrict__ temp,
                                                   const double* __restrict__ geom, double* __restrict__ grad, double* __restrict__ mu3, int* err,
                                                   double* __restrict__ src /* lctm2015 only */) {
  const Layout& L = P.L;
  // a warp covers cells i = 32 b - 15 .. 32 b + 16: cell 1 of a row starts a 128-byte line (ctx.hpp), so every row segment a warp
  // loads or stores is two whole lines (starting the warps at cell 0 made it three, two of them partial)
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) - G_ALIGN;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i < 0 || i > L.imx || j > L.jmx) return;
  const long long fs = L.fs, c = L.idx(i, j, k);
  double nan_probe = 0.0;
  double g[NG][3];
  green_gauss<NG>(P, q, temp, geom, c, g, nan_probe);
#pragma unroll
  for (int cc = 0; cc < NG; ++cc) {
#pragma unroll
    for (int d = 0; d < 3; ++d) grad[(3 * cc + d) * fs + c] = g[cc][d];
  }
  if (isnan(nan_probe)) atomicOr(err, F3D_ERR_NAN_GRADIENT);
  const double mu = molecular_viscosity(P, q, c, err);
  mu3[c] = mu;
  if (NG >= 5) {
    double mut, F1;
    eddy_viscosity<NG>(P, q, geom, c, g, mu, mut, F1);
    mu3[fs + c] = mut;
    if (NG >= 6) mu3[2 * fs + c] = F1;
    if constexpr (NG == 7) {
      if (src && i >= 1 && i <= L.imx - 1 && j >= 1 && j <= L.jmx - 1 && k >= 1 && k <= L.kmx - 1) {
        double Sk, Sw, Sg;
        lctm_sources(P, g, q[c], q[5 * fs + c], q[6 * fs + c], q[7 * fs + c], mu, mut, F1, mu3[3 * fs + c], geom[(long long)G_DIST * fs + c],
                     geom[(long long)G_VOL * fs + c], Sk, Sw, Sg);
        src[c] = Sk; src[fs + c] = Sw; src[2 * fs + c] = Sg;
      }
    }
  }
  if (NG == 5) {
    // Spalart-Allmaras: the cross-diffusion scalar CD2 = grad(rho) . grad(nu-tilde) of add_sa_source (source.f90:883-952), whose density
    // gradient is a Green-Gauss sum over the six neighbours.  Made here, where the gradient of nu-tilde is in registers, and carried to the
    // sweep in the 16th gradient field: in the sweep it sat on the I-row warps alone, which the other warps then waited for.
    // KEPT DEFECT: the normal of the low K face is (nx, nx, nx) (source.f90:901).
    const long long sj = L.sj, sk = L.sk;
    const double* __restrict__ gI = geom + (long long)G_IA * fs;
    const double* __restrict__ gJ = geom + (long long)G_JA * fs;
    const double* __restrict__ gK = geom + (long long)G_KA * fs;
    const double density = q[c];
    const double RhoFace[6] = {q[c - 1] + density, q[c - sj] + density, q[c - sk] + density, q[c + 1] + density, q[c + sj] + density, q[c + sk] + density};
    const double volc = geom[(long long)G_VOL * fs + c];
    double gradrho[3];
#pragma unroll
    for (int dd = 0; dd < 3; ++dd) {
      const double n0 = gI[(1 + dd) * fs + c], n1 = gJ[(1 + dd) * fs + c], n2 = gK[fs + c];
      const double n3 = gI[(1 + dd) * fs + c + 1], n4 = gJ[(1 + dd) * fs + c + sj], n5 = gK[(1 + dd) * fs + c + sk];
      gradrho[dd] = (-(RhoFace[0]) * n0 * gI[c] - (RhoFace[1]) * n1 * gJ[c] - (RhoFace[2]) * n2 * gK[c] + (RhoFace[3]) * n3 * gI[c + 1] +
                     (RhoFace[4]) * n4 * gJ[c + sj] + (RhoFace[5]) * n5 * gK[c + sk]) / (2.0 * volc);
    }
    grad[15 * fs + c] = ((gradrho[0] * g[4][0]) + (gradrho[1] * g[4][1]) + (gradrho[2] * g[4][2]));
  }
}

// Ghost-gradient rule + ghost mu_t / F1 on the physical faces (gradients.f90:486-676, viscosity.f90:165-212, 408-465, 488-531).  The
// reference applies the rule BEFORE it computes the viscosities and then copies mu_t / F1 from the interior cell on most boundary types;
// where it does not copy (periodic interfaces -10, total pressure -11) the ghost values follow from the ghost cell's own state and its
// rule-made gradients, recomputed here.  A separate kernel on purpose: folded into k_gradients (the ghost thread evaluating the interior
// cell itself) the pair measured 1.88-2.05 ms against 1.43 + 0.23 ms at 256^3 (divergent boundary warps, 26 more registers:
// profiles/r02_summary.md).
template <int NG>
__global__ void k_gradient_bc(const Params P, const double* __restrict__ q, const double* __restrict__ temp, const double* __restrict__ geom,
                              double* __restrict__ grad, double* __restrict__ mu3, const double* __restrict__ rec_all, const long long* rec_off,
                              int face_mask) {
  // one launch for all physical faces (blockIdx.z = face-1): each face reads interior gradients and writes its own ghost cells
  const int face = blockIdx.z + 1;
  if (!(face_mask >> (face - 1) & 1)) return;
  const double* __restrict__ rec = rec_all + rec_off[face - 1];
  const Layout& L = P.L;
  const int ax = (face - 1) / 2;
  const bool lo = (face % 2) == 1;
  const int a_ax = (ax == 0) ? 1 : 0, b_ax = (ax == 2) ? 1 : 2;
  const int mx[3] = {L.imx, L.jmx, L.kmx};
  const long long st[3] = {1, L.sj, L.sk};
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y * blockDim.y + threadIdx.y;
  const int na = mx[a_ax] - 1, nb = mx[b_ax] - 1;
  if (a >= na || b >= nb) return;
  int idx[3]; idx[a_ax] = a + 1; idx[b_ax] = b + 1; idx[ax] = lo ? 1 : mx[ax] - 1;
  const long long ci = L.idx(idx[0], idx[1], idx[2]);       // interior cell
  const long long cg = lo ? ci - st[ax] : ci + st[ax];       // ghost cell
  const long long fs = L.fs;
  const double* r = rec + 4 * ((long long)b * na + a);
  const double A = r[0], nx = r[1], ny = r[2], nz = r[3];
  const double vol = geom[(long long)G_VOL * fs + ci];
  const double c_x = A * nx / vol, c_y = A * ny / vol, c_z = A * nz / vol;
  const double sig = lo ? 1.0 : -1.0;
  const int id = P.bc_id[face - 1];
  const double ft = P.fixed[F3D_FIX_WALL_TEMP][face - 1];
  // all loads first, then all stores: grad is read and written here, so stores between the loads would serialise the NG
  // components into NG dependent round trips to HBM (on an i face every value is its own 32-byte sector)
  double qI[NG], qG[NG], gI[NG][3], gG[NG][3];
#pragma unroll
  for (int cc = 0; cc < NG; ++cc) {
    // slot cc holds variable cc+2 of qp(2:n_var): u,v,w,p,[k,omega]; slot 4 (cc == 3) is then overwritten with T
    qI[cc] = (cc == 3) ? temp[ci] : q[(long long)(cc + 1) * fs + ci];
    qG[cc] = (cc == 3) ? temp[cg] : q[(long long)(cc + 1) * fs + cg];
    gI[cc][0] = grad[(3 * cc + 0) * fs + ci]; gI[cc][1] = grad[(3 * cc + 1) * fs + ci]; gI[cc][2] = grad[(3 * cc + 2) * fs + ci];
  }
#pragma unroll
  for (int cc = 0; cc < NG; ++cc) {
    const double gIx = gI[cc][0], gIy = gI[cc][1], gIz = gI[cc][2];
    double gx = sig * (qI[cc] - qG[cc]) * c_x, gy = sig * (qI[cc] - qG[cc]) * c_y, gz = sig * (qI[cc] - qG[cc]) * c_z;
    if (cc == 3 && id == -5 && (ft < 1. && ft >= 0.)) { gx = -gIx; gy = -gIy; gz = -gIz; }   // adiabatic wall
    const double dot = (gIx * nx) + (gIy * ny) + (gIz * nz);
    gG[cc][0] = gx + (gIx - dot * nx);
    gG[cc][1] = gy + (gIy - dot * ny);
    gG[cc][2] = gz + (gIz - dot * nz);
    grad[(3 * cc + 0) * fs + cg] = gG[cc][0];
    grad[(3 * cc + 1) * fs + cg] = gG[cc][1];
    grad[(3 * cc + 2) * fs + cg] = gG[cc][2];
  }
  if (NG >= 5) {
    // boundary types that copy the interior mu_t (and F1): sa and sst -4..-1, -6..-9; kkl the same without the pole -7; the wall takes -mu_t
    const bool listed = id == -1 || id == -2 || id == -3 || id == -4 || id == -6 || id == -8 || id == -9 || (id == -7 && !(NG >= 6 && P.kkl));
    if (id == -5 || listed) {
      mu3[fs + cg] = (id == -5) ? -mu3[fs + ci] : mu3[fs + ci];
      if (NG >= 6 && !P.kkl) mu3[2 * fs + cg] = mu3[2 * fs + ci];
    } else {
      double mut, F1;
      eddy_viscosity<NG>(P, q, geom, cg, gG, mu3[cg], mut, F1);
      mu3[fs + cg] = mut;
      if (NG >= 6) mu3[2 * fs + cg] = F1;
    }
  }
}

// k-kL: magnitude of the second velocity derivatives the von Karman length scale needs (source.f90:700-760): for each velocity component the
// sum over the three directions of the Green-Gauss derivative of its first derivative, from the finished gradient arrays (ghost rule applied).
// Written into aux field 2 (the F1 slot, unused by k-kL) of the interior cells, where the sweep stages it.
__global__ void __launch_bounds__(128) k_kkl_udd(const Params P, const double* __restrict__ geom, const double* __restrict__ grad, double* __restrict__ mu3) {
  const Layout& L = P.L;
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y * blockDim.y + threadIdx.y, k = 1 + blockIdx.z;
  if (i > L.imx - 1 || j > L.jmx - 1) return;
  const long long fs = L.fs, c = L.idx(i, j, k), sj = L.sj, sk = L.sk;
  const double* __restrict__ gI = geom + (long long)G_IA * fs;
  const double* __restrict__ gJ = geom + (long long)G_JA * fs;
  const double* __restrict__ gK = geom + (long long)G_KA * fs;
  const double volc = geom[(long long)G_VOL * fs + c];
  double lap[3] = {0., 0., 0.};
#pragma unroll
  for (int dd = 0; dd < 3; ++dd) {
    const double wIl = gI[(1 + dd) * fs + c] * gI[c], wIh = gI[(1 + dd) * fs + c + 1] * gI[c + 1];
    const double wJl = gJ[(1 + dd) * fs + c] * gJ[c], wJh = gJ[(1 + dd) * fs + c + sj] * gJ[c + sj];
    const double wKl = gK[(1 + dd) * fs + c] * gK[c], wKh = gK[(1 + dd) * fs + c + sk] * gK[c + sk];
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
      const double* __restrict__ gf = grad + (long long)(3 * cc + dd) * fs;
      const double g0 = gf[c];
      const double s2 = (-(gf[c - 1] + g0) * wIl - (gf[c - sj] + g0) * wJl - (gf[c - sk] + g0) * wKl + (gf[c + 1] + g0) * wIh + (gf[c + sj] + g0) * wJh +
                         (gf[c + sk] + g0) * wKh) / (2 * volc);
      lap[cc] += s2;
    }
  }
  mu3[2 * fs + c] = sqrt(lap[0] * lap[0] + lap[1] * lap[1] + lap[2] * lap[2]);
}

// lctm2015, CC.f90:73-122: find_CCnormal = Green-Gauss gradient g of the wall distance on cells 0..imx (compute_gradient :125-200),
// normalised with |g| + 1e-12; find_DCCVn then -- KEPT DEFECT -- differentiates `dist` once more instead of CCVn = CCnormal . velocity,
// so the "wall-normal velocity gradient" of add_sst_source_lctm2015 is DCCVn . CCnormal = |g|^2 / (|g| + 1e-12): a field fixed by the
// grid.  Computed once per geometry / wall distance (api.cu:init_aux_fields) into aux field 3, where the sweep stages it.
__global__ void k_dvdy(const Params P, const double* __restrict__ geom, double* __restrict__ out) {
  const Layout& L = P.L;
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y, k = blockIdx.z;
  if (i > L.imx || j > L.jmx) return;
  const long long fs = L.fs, c = L.idx(i, j, k), sj = L.sj, sk = L.sk;
  const double* gI = geom + (long long)G_IA * fs;
  const double* gJ = geom + (long long)G_JA * fs;
  const double* gK = geom + (long long)G_KA * fs;
  const double* var = geom + (long long)G_DIST * fs;
  const double v0 = var[c];
  double g[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    g[d] = (-(var[c - 1] + v0) * gI[(1 + d) * fs + c] * gI[c] - (var[c - sj] + v0) * gJ[(1 + d) * fs + c] * gJ[c] - (var[c - sk] + v0) * gK[(1 + d) * fs + c] * gK[c] +
            (var[c + 1] + v0) * gI[(1 + d) * fs + c + 1] * gI[c + 1] + (var[c + sj] + v0) * gJ[(1 + d) * fs + c + sj] * gJ[c + sj] +
            (var[c + sk] + v0) * gK[(1 + d) * fs + c + sk] * gK[c + sk]) /
           (2 * geom[(long long)G_VOL * fs + c]);
  }
  const double mag = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
  const double nx = g[0] / (mag + 1e-12), ny = g[1] / (mag + 1e-12), nz = g[2] / (mag + 1e-12);
  out[c] = g[0] * nx + g[1] * ny + g[2] * nz;
}

int launch_dvdy(Ctx* ctx) {
  const Layout& L = ctx->P.L;
  dim3 block(32, 4, 1), grid((L.imx + 1 + 31) / 32, (L.jmx + 1 + 3) / 4, L.kmx + 1);
  k_dvdy<<<grid, block, 0, ctx->stream>>>(ctx->P, ctx->geom, ctx->mu + 3 * L.fs);
  F3D_CUDA(cudaGetLastError());
  return 0;
}

int launch_gradients(Ctx* ctx) {
  const Layout& L = ctx->P.L;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->timing) {
    if (ctx->ev_used2 == ctx->ev_pool2.size()) {
      cudaEvent_t x, y; cudaEventCreate(&x); cudaEventCreate(&y);
      ctx->ev_pool2.emplace_back(x, y);
    }
    e0 = ctx->ev_pool2[ctx->ev_used2].first; e1 = ctx->ev_pool2[ctx->ev_used2].second; ctx->ev_used2++;
    cudaEventRecord(e0, ctx->stream);
  }
  dim3 block(32, 4, 1);
  dim3 grid((L.imx + 1 + G_ALIGN + 31) / 32, (L.jmx + 1 + 3) / 4, L.kmx + 1);
#define F3D_GRAD_LAUNCH(NG_) k_gradients<NG_><<<grid, block, 0, ctx->stream>>>(ctx->P, ctx->qp, ctx->temp, ctx->geom, ctx->grad, ctx->mu, ctx->err_dev, ctx->src)
  if (ctx->P.sa) F3D_GRAD_LAUNCH(5);
  else if (ctx->P.lctm) F3D_GRAD_LAUNCH(7);
  else if (ctx->P.sst) F3D_GRAD_LAUNCH(6);
  else F3D_GRAD_LAUNCH(4);
#undef F3D_GRAD_LAUNCH
  ctx->launches++;
  const int mx[3] = {L.imx, L.jmx, L.kmx};
  int mask = 0, na = 1, nb = 1;
  for (int face = 1; face <= 6; ++face) {
    if (ctx->P.bc_id[face - 1] >= 0) continue;   // "if (bc%imin_id < 0)" -- includes -10
    const int ax = (face - 1) / 2;
    const int a_ax = (ax == 0) ? 1 : 0, b_ax = (ax == 2) ? 1 : 2;
    mask |= 1 << (face - 1);
    na = std::max(na, mx[a_ax] - 1); nb = std::max(nb, mx[b_ax] - 1);
  }
  if (mask) {
    dim3 g2((na + 31) / 32, (nb + 3) / 4, 6);
#define F3D_GBC_LAUNCH(NG_) k_gradient_bc<NG_><<<g2, block, 0, ctx->stream>>>(ctx->P, ctx->qp, ctx->temp, ctx->geom, ctx->grad, ctx->mu, ctx->gbc, ctx->gbc_off_dev, mask)
    if (ctx->P.sa) F3D_GBC_LAUNCH(5);
    else if (ctx->P.lctm) F3D_GBC_LAUNCH(7);
    else if (ctx->P.sst) F3D_GBC_LAUNCH(6);
    else F3D_GBC_LAUNCH(4);
#undef F3D_GBC_LAUNCH
    ctx->launches++;
  }
  if (ctx->P.kkl) {
    k_kkl_udd<<<dim3((L.imx - 1 + 31) / 32, (L.jmx - 1 + 3) / 4, L.kmx - 1), block, 0, ctx->stream>>>(ctx->P, ctx->geom, ctx->grad, ctx->mu);
    ctx->launches++;
  }
  if (ctx->timing) cudaEventRecord(e1, ctx->stream);
  F3D_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace f3d
