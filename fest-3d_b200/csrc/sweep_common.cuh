// Device helpers shared by the generations of the fused sweep kernel (sweep.cu: generation 2, sweep3.cu: generation 3):
// kernel arguments, the field order of a staged cell record, face reconstruction from staged values, the viscous face flux
// and the per-face evaluation.  Tile-dependent strides (PS: slots of a staged record plane, PSQ: slots of a staged q plane)
// are template parameters.
#pragma once
#include "ctx.hpp"
#include "physics.cuh"

namespace f3d {

template <int NV, bool VISC>
struct RecF {   // fields of the non-q part of a cell record
  static constexpr bool SST = (NV >= 7);                  // sst / sst2003 (k, omega), k-kL (k, kL)
  static constexpr bool LCTM = (NV == 8);                  // sst / sst2003 + the intermittency of transition = lctm2015 (variable 8)
  static constexpr bool SA = (NV == 6);                    // Spalart-Allmaras (nu-tilde)
  // staged gradient components: u, v, w, T [, k, omega | nu-tilde].  The intermittency gradient (component 6 of lctm2015) is NOT staged:
  // with eight variables the shared memory of an SM is full, so its diffusion flux reads it -- and the cell centres -- from global memory
  static constexpr int NG = SST ? 6 : (SA ? 5 : 4);
  static constexpr int NGF = VISC ? 3 * NG : 0;            // gradient component c, direction d -> field 3*c+d
  static constexpr int NGFS = (NGF + 1) & ~1;              // staged slots of the gradient fields: an even count (tensor-map boxes
                                                           // need 128-byte aligned sub-boxes; sa has 15 fields -> one unused slot)
  static constexpr int NMU = VISC ? (SST ? 3 : (SA ? 2 : 1)) : 0;   // mu, mu_t, F1
  static constexpr int OFF_MU = NGFS, OFF_C = NGFS + NMU;  // then the cell centre x,y,z (lctm2015: the CC.f90 field instead, grad.cu:k_dvdy)
  static constexpr int NAUX = VISC ? (LCTM ? NMU + 1 : NMU + 3) : 0;   // "aux" fields behind the gradients: mu [, mu_t [, F1]], centre x,y,z
  static constexpr int NAUXS = (NAUX + 1) & ~1;
  static constexpr int NR = VISC ? NGFS + NAUXS : 0;
};

__device__ __forceinline__ void flag_error(int* err, int cls, int i, int j, int k) {
  int old = atomicOr(&err[0], cls);
  if ((old & cls) == 0) { err[1] = i; err[2] = j; err[3] = k; }
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct KArgs {
  const double* __restrict__ q;       // nv fields, ghost-filled
  const double* __restrict__ quse;    // U_store or q
  double* __restrict__ qnew;          // nv fields (update mode)
  double* __restrict__ residue;       // nv fields (residue mode)
  double* __restrict__ rstore;        // nv fields or nullptr
  double* __restrict__ dt;            // 1 field
  const double* __restrict__ geom;
  const double* __restrict__ grad;    // staged-gradient path only (sweep3_kernel.cuh)
  const double* __restrict__ mu;      //   mu, mu_t, F1 + centre copy
  const double* __restrict__ src;     //   lctm2015: S_k V, S_omega V, S_gamma V of every interior cell (grad.cu:lctm_sources)
  const double* __restrict__ gbc;     // face records (A, nx, ny, nz) of the ghost-gradient rule, six faces back to back
  long long gbc_off[6];               // offset of every face's records inside gbc
  double* __restrict__ red;           // per-CTA partials [(nv+1) * n_cta]
  int* err;
  int mode, first_stage, want_norms, have_store, use_store_sum, kchunk;
  double TF, SF;
};

// Pressure-based switching (muscl.f90:37-112, ppm.f90:108-170): weight of the cell at position pos along a direction with node
// count mx.  Interior cells take it from their two neighbours; the ghost positions 0 and mx copy the value of the first / last
// interior cell, i.e. they are built from p(2), p(0) and p(mx), p(mx-2): p_far is the pressure two cells inwards.
__device__ __forceinline__ double pb_pdif(const Params& P, double p_m, double p_0, double p_p, double p_far, int pos, int mx) {
  const double a = (pos == 0) ? p_far : ((pos == mx) ? p_0 : p_p);
  const double b = (pos == 0) ? p_0 : ((pos == mx) ? p_far : p_m);
  const double pd2 = fabs(a - b);
  return 1 - (pd2 / (pd2 + P.pressure_inf));
}

// Values a cell contributes to its two faces along one direction, all variables.  `pos` is the cell's index along the
// direction; the first / last interior cell next to a physical boundary is re-done with the boundary formula when
// ppm_flag is set (boundary_state_reconstruction.f90:93-123).
template <int NV, int INTERP, bool PB = false>
__device__ __forceinline__ void line_cell_values(const Params& P, const double* __restrict__ q, const double* __restrict__ vol,
                                                 long long c, long long s, int pos, int mx, int dir, double (&to_hi)[NV], double (&to_lo)[NV]) {
  const bool redo = (INTERP != F3D_INTERP_NONE) && P.ppm_flag && ((pos == 1 && P.phys[2 * dir]) || (pos == mx - 1 && P.phys[2 * dir + 1]));
  double vl[7];
  if (INTERP == F3D_WENO_NM) {
#pragma unroll
    for (int m = 1; m <= 5; ++m) vl[m] = vol[c + (m - 3) * s];
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const double* __restrict__ qv = q + (long long)v * P.L.fs;
    const int lim = (v >= 5) ? P.tlimiter[dir] : P.limiter[dir];
    if (redo) {
      boundary_cell_face_values(qv[c - s], qv[c], qv[c + s], lim, to_hi[v], to_lo[v]);
    } else {
      double ql[7];
      if (INTERP == F3D_INTERP_NONE) { ql[3] = qv[c]; }
      else if (INTERP == F3D_MUSCL) { ql[2] = qv[c - s]; ql[3] = qv[c]; ql[4] = qv[c + s]; }
      else {
#pragma unroll
        for (int m = 1; m <= 5; ++m) ql[m] = qv[c + (m - 3) * s];
      }
      cell_face_values<INTERP>(ql, vl, lim, to_hi[v], to_lo[v]);
    }
  }
  if (PB && (INTERP == F3D_PPM || INTERP == F3D_MUSCL) && !redo && P.pb_switch[dir]) {
    const double* __restrict__ pv = q + 4LL * P.L.fs;
    const double pd = pb_pdif(P, pv[c - s], pv[c], pv[c + s], (pos == 0) ? pv[c + 2 * s] : ((pos == mx) ? pv[c - 2 * s] : 0.0), pos, mx);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const double q0 = q[(long long)v * P.L.fs + c];
      to_hi[v] = q0 + (pd * (to_hi[v] - q0));
      to_lo[v] = q0 - (pd * (q0 - to_lo[v]));
    }
  }
}


// Koren limiter psi(r) = max(0, min(2r, (2/3)(r-1) + 1, 2)) (muscl.f90:176-181) as ONE fused multiply-add psi = s*r+ + t whose
// coefficients are picked by the range r falls in: r <= 0: r+ = 0 -> 0;  0 < r < 1/4: 2 r;  1/4 <= r < 5/2: (2/3) r + 1/3;
// r >= 5/2: 2.  The break points 1/4 and 5/2 are where the reference's min() changes its argument, and both have a zero low
// word, so a signed compare of the HIGH word of r decides the range exactly: the three FP64 compares (DSETP, on the FP64 pipe,
// each with the NaN-propagation select sequence the compiler emits for min/max) become integer compares, and the 2r, r-1 and
// (2/3)(.)+1 operations one DFMA: 6 -> 1 FP64 instructions per psi, 42 psi per cell and direction.  (2/3) r + 1/3 differs from
// (2/3)(r-1) + 1 by at most one rounding (2e-16 relative on psi); F3D_KOREN_REF keeps the reference's operation sequence.
__device__ __forceinline__ double koren_psi(double r) {
#ifdef F3D_KOREN_REF
  return dmax(0., dmin(dmin(2 * r, (2. / 3.) * (r - 1.0) + 1.0), 2.));
#else
  const int h = __double2hiint(r);
  const bool high = h >= 0x40040000;                              // r >= 2.5 (a negative r has h < 0)
  const bool mid = (h >= 0x3FD00000) && !high;                    // 0.25 <= r < 2.5
  const double rp = __hiloint2double(max(h, 0), __double2loint(r));   // r <= 0 -> a value below 2^-1022*2^20: psi = 0 to 1e-300
  const int lo = mid ? 0x55555555 : 0;
  const int s_hi = mid ? 0x3FE55555 : (high ? 0 : 0x40000000);     // 2/3 | 0 | 2
  const int t_hi = mid ? 0x3FD55555 : (high ? 0x40000000 : 0);     // 1/3 | 2 | 0
  return fma(__hiloint2double(s_hi, lo), rp, __hiloint2double(t_hi, lo));
#endif
}

// Koren-limited kappa = 1/3 MUSCL values of all variables of one cell (muscl.f90:161-196): ONE branch-free loop, so that the NV
// independent dependency chains interleave (a DFMA has 8.4 cycles of latency and the pipe takes one every 2.1:
// profiles/r01_fp64_ops_microbench.txt) and the code exists once.  The limiter switches (0 / 1 per direction, flow and turbulence
// variables separately) select between psi and 1: the reference's 1 - (1 - psi)*switch is psi to within an ulp for switch = 1 and
// exactly 1 for switch = 0.
template <int NV>
__device__ __forceinline__ void muscl_all(const double (&qm)[NV], const double (&q0)[NV], const double (&qp)[NV], int lim, int tlim, double (&to_hi)[NV],
                                          double (&to_lo)[NV]) {
  // q0 +- 0.25*((1 -+ kappa) psi1 bd + (1 +- kappa) psi2 fd) with the constant factors folded: ca = 0.25 (1 - kappa),
  // cb = 0.25 (1 + kappa), and the two limited differences p1 = psi1 bd, p2 = psi2 fd shared by both faces
  const double kappa = 1. / 3.;
  const double ca = 0.25 * (1. - kappa), cb = 0.25 * (1. + kappa);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const bool on = ((v >= 5) ? tlim : lim) != 0;
    const double fd = qp[v] - q0[v], bd = q0[v] - qm[v];
    // x + sign(1e-14, x) written as sign(|x| + 1e-14, x): the same value bit for bit, but the constant comes from the constant
    // bank as an operand instead of being built in two registers per use
    double psi1 = koren_psi(fd * rcp64(copysign(fabs(bd) + 1e-14, bd)));
    double psi2 = koren_psi(bd * rcp64(copysign(fabs(fd) + 1e-14, fd)));
    psi1 = on ? psi1 : 1.0;
    psi2 = on ? psi2 : 1.0;
    const double p1 = psi1 * bd, p2 = psi2 * fd;
    to_hi[v] = fma(ca, p1, fma(cb, p2, q0[v]));
    to_lo[v] = fma(-cb, p1, fma(-ca, p2, q0[v]));
  }
}

// MUSCL / first-order values of one cell along one direction from three staged values per variable
// (muscl.f90:161-196; boundary_state_reconstruction.f90:93-123 for the first / last interior cell when ppm_flag is set: done as an
// overwrite AFTER the regular formula, so that the hot path carries no merge of the two)
template <int NV, int INTERP, bool PB = false>
__device__ __forceinline__ void recon3(const Params& P, const double (&qm)[NV], const double (&q0)[NV], const double (&qp)[NV], int pos, int mx,
                                       int dir, double (&to_hi)[NV], double (&to_lo)[NV], double p_far = 0.0) {
  if (INTERP == F3D_INTERP_NONE) {
#pragma unroll
    for (int v = 0; v < NV; ++v) { to_hi[v] = q0[v]; to_lo[v] = q0[v]; }
    return;
  }
  const int lim = P.limiter[dir], tlim = P.tlimiter[dir];
  muscl_all<NV>(qm, q0, qp, lim, tlim, to_hi, to_lo);
  if (PB && P.pb_switch[dir]) {   // pressure-based switching (muscl.f90:231-243): both face values are pulled towards the cell value
    const double pd = pb_pdif(P, qm[4], q0[4], qp[4], p_far, pos, mx);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      to_hi[v] = q0[v] + (pd * (to_hi[v] - q0[v]));
      to_lo[v] = q0[v] - (pd * (q0[v] - to_lo[v]));
    }
  }
  const bool redo = P.ppm_flag && ((pos == 1 && P.phys[2 * dir]) || (pos == mx - 1 && P.phys[2 * dir + 1]));
  if (redo) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      double h, l;
      boundary_cell_face_values(qm[v], q0[v], qp[v], (v >= 5) ? tlim : lim, h, l);
      to_hi[v] = h; to_lo[v] = l;
    }
  }
}

// F <- (F - laminar) - sst for the face between cells lo and hi (viscous.f90:209-323, 378-446); also the face terms
// A*mu/(rho*|dr.n|), A*mu_t/(rho*|dr.n|) of the viscous / turbulent time-step corrections (time.f90:396-421, 479-504:
// both cells that share a face use the mu and density of the cell on its high side).  ql/qh: staged q of the two cells
// (field stride PSQ), rl/rh: their staged records (field stride PS).
template <int NV, int PS, int PSQ>
__device__ __forceinline__ void viscous_face(const Params& P, const double* __restrict__ ql_, const double* __restrict__ qh_,
                                             const double* __restrict__ rl, const double* __restrict__ rh, double A, double nx, double ny,
                                             double nz, bool sst_on, bool need_dt, double (&F)[NV], double& vis, double& tur, bool kkl = false,
                                             const KArgs* a = nullptr, long long fs = 0, long long cgl = 0, long long cgh = 0) {
  using R = RecF<NV, true>;
  constexpr bool SST = R::SST, SA = R::SA, LCTM = R::LCTM, TURB = SST || SA;
  constexpr int NG = R::NG;
  double dx, dy, dz;
  if (LCTM) {   // the centres are not staged (RecF): cells cgl / cgh of the global arrays
    const double* __restrict__ cx = a->geom + (long long)G_CX * fs;
    dx = cx[cgh] - cx[cgl]; dy = cx[fs + cgh] - cx[fs + cgl]; dz = cx[2 * fs + cgh] - cx[2 * fs + cgl];
  } else {
    dx = rh[(R::OFF_C + 0) * PS] - rl[(R::OFF_C + 0) * PS]; dy = rh[(R::OFF_C + 1) * PS] - rl[(R::OFF_C + 1) * PS];
    dz = rh[(R::OFF_C + 2) * PS] - rl[(R::OFF_C + 2) * PS];
  }
  const double inv_d = rsqrt64(dx * dx + dy * dy + dz * dz);   // 1 / d_LR
  const double ex = dx * inv_d, ey = dy * inv_d, ez = dz * inv_d;
  double ql[NV], qh[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) { ql[v] = ql_[v * PSQ]; qh[v] = qh_[v * PSQ]; }
  double del[NG];
  del[0] = qh[1] - ql[1]; del[1] = qh[2] - ql[2]; del[2] = qh[3] - ql[3];
  {
    const double T_LE = ql[4] * rcp64(ql[0] * P.R_gas), T_RE = qh[4] * rcp64(qh[0] * P.R_gas);
    del[3] = T_RE - T_LE;
  }
  if (SST) { del[4] = qh[5] - ql[5]; del[5] = qh[6] - ql[6]; }
  if (SA) del[4] = qh[5] - ql[5];
#ifdef F3D_VISC_HALVES   // the reference's operation sequence: every face average formed with its own multiplication by 0.5
  double G[NG][3];
#pragma unroll
  for (int c = 0; c < NG; ++c) {
    const double ax = 0.5 * (rl[(3 * c) * PS] + rh[(3 * c) * PS]), ay = 0.5 * (rl[(3 * c + 1) * PS] + rh[(3 * c + 1) * PS]),
                 az = 0.5 * (rl[(3 * c + 2) * PS] + rh[(3 * c + 2) * PS]);
    const double nc = (del[c] - (ax * dx + ay * dy + az * dz)) * inv_d;
    G[c][0] = ax + (nc * ex);
    G[c][1] = ay + (nc * ey);
    G[c][2] = az + (nc * ez);
  }
  const double mu_hi = rh[R::OFF_MU * PS];
  const double mu_f = 0.5 * (rl[R::OFF_MU * PS] + mu_hi);
  const double mut_hi = TURB ? rh[(R::OFF_MU + 1) * PS] : 0.0;
  const double mut_f = TURB ? 0.5 * (rl[(R::OFF_MU + 1) * PS] + mut_hi) : 0.0;
  const double tmu = mu_f + mut_f;
  const double div3 = (G[0][0] + G[1][1] + G[2][2]) * (1. / 3.);
  const double Txx = 2. * tmu * (G[0][0] - div3), Tyy = 2. * tmu * (G[1][1] - div3), Tzz = 2. * tmu * (G[2][2] - div3);
  const double Txy = tmu * (G[1][0] + G[0][1]), Txz = tmu * (G[2][0] + G[0][2]), Tyz = tmu * (G[2][1] + G[1][2]);
  const double Kh = (mu_f * P.inv_Pr + mut_f * P.inv_tPr) * P.gm * P.R_gas * P.inv_gm1;
  const double Qx = Kh * G[3][0], Qy = Kh * G[3][1], Qz = Kh * G[3][2];
  const double A4 = A, mu_s = mu_f, mut_s = mut_f;
#else
  // Every face average 0.5*(l + h) of the reference enters the flux linearly, and scaling by a power of two commutes with
  // rounding: the SUMS are carried instead (G2 = 2 G, mu_s = 2 mu_f, mut_s = 2 mu_t,f) and the factors collected into the few
  // coefficients that multiply them (tmu, tmu/2, Kh/2, A/4) -- the same bits as the reference's sequence with 17 fewer
  // multiplications per face (F3D_VISC_HALVES keeps that sequence).
  double G[NG][3];   // 2 x the face gradient
#pragma unroll
  for (int c = 0; c < NG; ++c) {
    const double ax = rl[(3 * c) * PS] + rh[(3 * c) * PS], ay = rl[(3 * c + 1) * PS] + rh[(3 * c + 1) * PS], az = rl[(3 * c + 2) * PS] + rh[(3 * c + 2) * PS];
    const double nc = ((2. * del[c]) - (ax * dx + ay * dy + az * dz)) * inv_d;
    G[c][0] = ax + (nc * ex);
    G[c][1] = ay + (nc * ey);
    G[c][2] = az + (nc * ez);
  }
  const double mu_hi = rh[R::OFF_MU * PS];
  const double mu_s = rl[R::OFF_MU * PS] + mu_hi;
  const double mut_hi = TURB ? rh[(R::OFF_MU + 1) * PS] : 0.0;
  const double mut_s = TURB ? rl[(R::OFF_MU + 1) * PS] + mut_hi : 0.0;
  const double tmu2 = mu_s + mut_s;
  const double tmu = 0.5 * tmu2, tmuh = 0.25 * tmu2;
  const double div3 = (G[0][0] + G[1][1] + G[2][2]) * (1. / 3.);
  const double Txx = tmu * (G[0][0] - div3), Tyy = tmu * (G[1][1] - div3), Tzz = tmu * (G[2][2] - div3);
  const double Txy = tmuh * (G[1][0] + G[0][1]), Txz = tmuh * (G[2][0] + G[0][2]), Tyz = tmuh * (G[2][1] + G[1][2]);
  const double Kh = 0.25 * ((mu_s * P.inv_Pr + mut_s * P.inv_tPr) * P.gm * P.R_gas * P.inv_gm1);
  const double Qx = Kh * G[3][0], Qy = Kh * G[3][1], Qz = Kh * G[3][2];
  const double A4 = 0.25 * A;
#endif
  const double uf = 0.5 * (ql[1] + qh[1]), vf = 0.5 * (ql[2] + qh[2]), wf = 0.5 * (ql[3] + qh[3]);
  F[1] = F[1] - ((Txx * nx + Txy * ny + Txz * nz) * A);
  F[2] = F[2] - ((Txy * nx + Tyy * ny + Tyz * nz) * A);
  F[3] = F[3] - ((Txz * nx + Tyz * ny + Tzz * nz) * A);
  F[4] = F[4] - (A * (((Txx * uf + Txy * vf + Txz * wf + Qx) * nx) + ((Txy * uf + Tyy * vf + Tyz * wf + Qy) * ny) +
                      ((Txz * uf + Tyz * vf + Tzz * wf + Qz) * nz)));
  if (SST && (sst_on || kkl)) {   // k-kL (viscous.f90:450-567): the same form with sigma_k = sigma_phi = 1, in all three directions whatever kmx
    const double F1 = 0.5 * (rl[(R::OFF_MU + 2) * PS] + rh[(R::OFF_MU + 2) * PS]);
    const double sk = kkl ? 1.0 : kSigmaK1 * F1 + kSigmaK2 * (1.0 - F1);
    const double sw = kkl ? 1.0 : kSigmaW1 * F1 + kSigmaW2 * (1.0 - F1);
    const double rhof = 0.5 * (ql[0] + qh[0]);
    const double tkf = 0.5 * (ql[5] + qh[5]);
    const double Tk = -2.0 * rhof * tkf * (1. / 3.);
    const double dk = (A4 * ((mu_s + sk * mut_s) * (G[NG - 2][0] * nx + G[NG - 2][1] * ny + G[NG - 2][2] * nz)));
    const double dw = (A4 * ((mu_s + sw * mut_s) * (G[NG - 1][0] * nx + G[NG - 1][1] * ny + G[NG - 1][2] * nz)));
    F[1] = F[1] - (Tk * nx * A);
    F[2] = F[2] - (Tk * ny * A);
    F[3] = F[3] - (Tk * nz * A);
    F[4] = F[4] - dk;
    F[5] = F[5] - dk;
    F[6] = F[6] - dw;
  }
  if (LCTM && sst_on) {   // viscous.f90:659-746: diffusion of the intermittency with mu + mu_t; like the k / omega fluxes skipped on K faces when kmx == 2
    const double* __restrict__ gg = a->grad + 18 * fs;   // gradient component 6, directions x, y, z
    const double ax = gg[cgl] + gg[cgh], ay = gg[fs + cgl] + gg[fs + cgh], az = gg[2 * fs + cgl] + gg[2 * fs + cgh];
    const double nc = ((2. * (qh[7] - ql[7])) - (ax * dx + ay * dy + az * dz)) * inv_d;
    const double gx = ax + (nc * ex), gy = ay + (nc * ey), gz = az + (nc * ez);   // 2 x the face gradient
#ifdef F3D_VISC_HALVES
    F[7] = F[7] - (A * ((mu_s + mut_s) * ((0.5 * gx) * nx + (0.5 * gy) * ny + (0.5 * gz) * nz)));
#else
    F[7] = F[7] - (A4 * ((mu_s + mut_s) * (gx * nx + gy * ny + gz * nz)));
#endif
  }
  if (SA) {   // viscous.f90:570-656: its "mut_f" is rho_face * nu-tilde_face, not the eddy viscosity; K flux also when kmx == 2
    const double rhof = 0.5 * (ql[0] + qh[0]);
    const double mut_sa = 0.5 * (ql[5] + qh[5]) * rhof;
#ifdef F3D_VISC_HALVES
    F[5] = F[5] - (A * ((mu_s + mut_sa) * (G[4][0] * nx + G[4][1] * ny + G[4][2] * nz))) * (1.0 / kSigmaSA);
#else
    F[5] = F[5] - ((0.5 * A) * (((0.5 * mu_s) + mut_sa) * (G[4][0] * nx + G[4][1] * ny + G[4][2] * nz))) * (1.0 / kSigmaSA);
#endif
  }
  if (need_dt) {
    const double dn = fabs(((-dx) * nx) + ((-dy) * ny) + ((-dz) * nz));
    const double w = A * rcp64(qh[0] * dn);
    vis = w * mu_hi;
    if (TURB) tur = w * mut_hi;
  }
}

// One face: boundary overrides of the states (boundary_state_reconstruction.f90:124-131), inviscid flux times area
// (scheme.f90:68-109), viscous flux, and the face terms of the time step.  ql/qh, rl/rh: staged q / record of the cells on
// the low / high side; A, n = metrics of the face; f = the face index along direction d, m = node count along d.
template <int NV, int SCHEME, bool VISC, int PS, int PSQ>
__device__ __forceinline__ void face_eval(const Params& P, int d, const double* __restrict__ ql, const double* __restrict__ qh,
                                          const double* __restrict__ rl, const double* __restrict__ rh, double A, double nx, double ny,
                                          double nz, int f, int m, double (&L)[NV], double (&R)[NV], bool flux_on, bool need_dt,
                                          double (&F)[NV], double& lam, double& vis, double& tur, bool kkl = false, const KArgs* a = nullptr,
                                          long long cgl = 0, long long cgh = 0) {
  if (P.interpolant != F3D_INTERP_NONE) {
    if (f == 1 && P.phys[2 * d]) {
      const bool far = P.farlike[2 * d] != 0;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const double g = ql[v * PSQ], in = qh[v * PSQ];
        if (far) { L[v] = g; R[v] = g; } else { L[v] = 0.5 * (g + in); }
      }
    }
    if (f == m && P.phys[2 * d + 1]) {
      const bool far = P.farlike[2 * d + 1] != 0;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const double in = ql[v * PSQ], g = qh[v * PSQ];
        if (far) { L[v] = g; R[v] = g; } else { R[v] = 0.5 * (in + g); }
      }
    }
  }
  const double mask = (f == 1) ? P.zlo[d] : ((f == m) ? P.zhi[d] : 1.0);
  const double cbar = inviscid_flux<NV>(SCHEME >= 0 ? SCHEME : P.scheme, P.gm, P.MInf, L, R, A, nx, ny, nz, mask, flux_on, need_dt, F);
  if (need_dt) {   // time.f90:159-237: both cells of a face use the velocity of the cell on its high side
    const double vn = fabs((qh[1 * PSQ] * nx) + (qh[2 * PSQ] * ny) + (qh[3 * PSQ] * nz));
    lam = A * (vn + cbar);
  }
  if (VISC) viscous_face<NV, PS, PSQ>(P, ql, qh, rl, rh, A, nx, ny, nz, (NV >= 7) && flux_on, need_dt, F, vis, tur, kkl, a, P.L.fs, cgl, cgh);
}

}  // namespace f3d
