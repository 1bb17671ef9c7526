// Fused residual / time-step / update kernel, generation 1 ("plane sweep").
//
// One CTA owns a TX x TY column of cells and marches through a chunk of k planes.  Per plane every cell reconstructs its
// own face values ONCE per direction, every face flux (inviscid + viscous + the face terms of the local time step) is
// evaluated ONCE by the thread of the cell on its high side, and the results are exchanged through shared memory:
//
//   phase 1  reconstruct (cell -> value at its high face "hi", value at its low face "lo")        -> smem hi / lo
//   phase 2  face flux   (L = hi of the low neighbour, R = lo of this cell)                        -> smem F
//   phase 3  cell        residual = (F(i+1)-F(i)) + (G(j+1)-G(j)) + (H(k+1)-H(k)), SST source, local time step,
//                        point-implicit k/omega scaling, RK accumulate, conservative update, norms
//
// The k direction needs no neighbour thread: the hi value and the low-face flux of the previous plane stay in a
// thread-private ping-pong slot.  Cells just outside the tile in i and j are served by three extra "halo" warps
// (one for the two i columns, one for the low j row, one for the high j row), so no face is computed twice inside a
// tile and only (TX+1)/TX, (TY+1)/TY of the faces are computed twice across tiles.  The direction loop is NOT
// unrolled: one instance of the reconstruction and one of the flux code (the v0 kernel had 0.58 MB of SASS and was
// instruction-fetch bound, profiles/r01_v0_summary.md).  No face-state, flux or residual array reaches HBM on the
// update path; the residual-norm partials are reduced in-kernel (warp shuffle + block).
//
// Reference pipeline reproduced (src/update.f90:534-545, 228-491; src/face/state/*.f90;
// src/boundary/boundary_state_reconstruction.f90:93-131; src/face/flux/convective/*.f90 and scheme.f90:111-141;
// src/viscous.f90:144-447; src/source.f90:158-270; src/time.f90:122-246,366-531; src/resnorm.f90:171-199).
#include "ctx.hpp"
#include "physics.cuh"

namespace f3d {

constexpr int TX = 32, TY = 8;
constexpr int NMAIN = TX * TY;
constexpr int NT = NMAIN + 96;   // + i-halo warp, low-j-halo warp, high-j-halo warp

// shared-memory slot counts per direction (slots are [variable][slot], variable-major)
constexpr int SLOT_I = TY * (TX + 1);
constexpr int SLOT_J = (TY + 1) * TX;
constexpr int SLOT_K = 2 * NMAIN;
constexpr int SLOT_HF = SLOT_I + SLOT_J + SLOT_K;   // slots of one hi (or F) buffer family
constexpr int SLOT_LO = 3 * NMAIN + 64;             // private lo slots: main threads x 3 directions + the halo-high tasks

__host__ __device__ constexpr int sweep_smem_doubles(int nv) { return nv * SLOT_HF + nv * SLOT_LO + (nv + 3) * SLOT_HF; }

__device__ __forceinline__ void flag_error(int* err, int cls, int i, int j, int k) {
  int old = atomicOr(&err[0], cls);
  if ((old & cls) == 0) { err[1] = i; err[2] = j; err[3] = k; }
}

struct KArgs {
  const double* __restrict__ q;       // nv fields, ghost-filled
  const double* __restrict__ quse;    // U_store or q
  double* __restrict__ qnew;          // nv fields (update mode)
  double* __restrict__ residue;       // nv fields (residue mode)
  double* __restrict__ rstore;        // nv fields or nullptr
  double* __restrict__ dt;            // 1 field
  const double* __restrict__ geom;
  const double* __restrict__ grad;
  const double* __restrict__ mu;      // mu, mu_t, F1
  double* __restrict__ red;           // per-CTA partials [(nv+1) * n_cta]
  int* err;
  int mode, first_stage, want_norms, have_store, use_store_sum, kchunk;
  double TF, SF;
};

// Values a cell contributes to its two faces along one direction, all variables.  `pos` is the cell's index along the
// direction; the first / last interior cell next to a physical boundary is re-done with the boundary formula when
// ppm_flag is set (boundary_state_reconstruction.f90:93-123).
template <int NV, int INTERP>
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
}

// F <- (F - laminar) - sst for the face between cells lo and hi (viscous.f90:209-323, 378-446); also the face terms
// A*mu/(rho*|dr.n|), A*mu_t/(rho*|dr.n|) of the viscous / turbulent time-step corrections (time.f90:396-421, 479-504:
// both cells that share a face use the mu and density of the cell on its high side).
template <int NV>
__device__ __forceinline__ void viscous_face(const Params& P, const KArgs& a, long long lo, long long hi, double A, double nx, double ny,
                                             double nz, bool sst_on, bool need_dt, double (&F)[NV], double& vis, double& tur) {
  constexpr bool SST = (NV == 7);
  constexpr int NG = SST ? 6 : 4;
  const long long fs = P.L.fs;
  const double* __restrict__ q = a.q;
  const double* __restrict__ gc = a.geom + (long long)G_CX * fs;
  const double dx = gc[hi] - gc[lo], dy = gc[fs + hi] - gc[fs + lo], dz = gc[2 * fs + hi] - gc[2 * fs + lo];
  const double inv_d = rsqrt64(dx * dx + dy * dy + dz * dz);   // 1 / d_LR
  const double ex = dx * inv_d, ey = dy * inv_d, ez = dz * inv_d;
  double ql[NV], qh[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) { ql[v] = q[v * fs + lo]; qh[v] = q[v * fs + hi]; }
  double del[NG];
  del[0] = qh[1] - ql[1]; del[1] = qh[2] - ql[2]; del[2] = qh[3] - ql[3];
  {
    const double T_LE = ql[4] * rcp64(ql[0] * P.R_gas), T_RE = qh[4] * rcp64(qh[0] * P.R_gas);
    del[3] = T_RE - T_LE;
  }
  if (SST) { del[4] = qh[5] - ql[5]; del[5] = qh[6] - ql[6]; }
  double G[NG][3];
#pragma unroll
  for (int c = 0; c < NG; ++c) {
    const double* __restrict__ g0 = a.grad + (long long)(3 * c) * fs;
    const double ax = 0.5 * (g0[lo] + g0[hi]), ay = 0.5 * (g0[fs + lo] + g0[fs + hi]), az = 0.5 * (g0[2 * fs + lo] + g0[2 * fs + hi]);
    const double nc = (del[c] - (ax * dx + ay * dy + az * dz)) * inv_d;
    G[c][0] = ax + (nc * ex);
    G[c][1] = ay + (nc * ey);
    G[c][2] = az + (nc * ez);
  }
  const double mu_hi = a.mu[hi];
  const double mu_f = 0.5 * (a.mu[lo] + mu_hi);
  const double mut_hi = SST ? a.mu[fs + hi] : 0.0;
  const double mut_f = SST ? 0.5 * (a.mu[fs + lo] + mut_hi) : 0.0;
  const double tmu = mu_f + mut_f;
  const double div3 = (G[0][0] + G[1][1] + G[2][2]) * (1. / 3.);
  const double Txx = 2. * tmu * (G[0][0] - div3), Tyy = 2. * tmu * (G[1][1] - div3), Tzz = 2. * tmu * (G[2][2] - div3);
  const double Txy = tmu * (G[1][0] + G[0][1]), Txz = tmu * (G[2][0] + G[0][2]), Tyz = tmu * (G[2][1] + G[1][2]);
  const double Kh = (mu_f * P.inv_Pr + mut_f * P.inv_tPr) * P.gm * P.R_gas * P.inv_gm1;
  const double Qx = Kh * G[3][0], Qy = Kh * G[3][1], Qz = Kh * G[3][2];
  const double uf = 0.5 * (ql[1] + qh[1]), vf = 0.5 * (ql[2] + qh[2]), wf = 0.5 * (ql[3] + qh[3]);
  F[1] = F[1] - ((Txx * nx + Txy * ny + Txz * nz) * A);
  F[2] = F[2] - ((Txy * nx + Tyy * ny + Tyz * nz) * A);
  F[3] = F[3] - ((Txz * nx + Tyz * ny + Tzz * nz) * A);
  F[4] = F[4] - (A * (((Txx * uf + Txy * vf + Txz * wf + Qx) * nx) + ((Txy * uf + Tyy * vf + Tyz * wf + Qy) * ny) +
                      ((Txz * uf + Tyz * vf + Tzz * wf + Qz) * nz)));
  if (SST && sst_on) {
    const double F1 = 0.5 * (a.mu[2 * fs + lo] + a.mu[2 * fs + hi]);
    const double sk = kSigmaK1 * F1 + kSigmaK2 * (1.0 - F1);
    const double sw = kSigmaW1 * F1 + kSigmaW2 * (1.0 - F1);
    const double rhof = 0.5 * (ql[0] + qh[0]);
    const double tkf = 0.5 * (ql[NV - 2] + qh[NV - 2]);
    const double Tk = -2.0 * rhof * tkf * (1. / 3.);
    const double dk = (A * ((mu_f + sk * mut_f) * (G[NG - 2][0] * nx + G[NG - 2][1] * ny + G[NG - 2][2] * nz)));
    const double dw = (A * ((mu_f + sw * mut_f) * (G[NG - 1][0] * nx + G[NG - 1][1] * ny + G[NG - 1][2] * nz)));
    F[1] = F[1] - (Tk * nx * A);
    F[2] = F[2] - (Tk * ny * A);
    F[3] = F[3] - (Tk * nz * A);
    F[4] = F[4] - dk;
    F[NV - 2] = F[NV - 2] - dk;
    F[NV - 1] = F[NV - 1] - dw;
  }
  if (need_dt) {
    const double dn = fabs(((-dx) * nx) + ((-dy) * ny) + ((-dz) * nz));
    const double w = A * rcp64(qh[0] * dn);
    vis = w * mu_hi;
    if (SST) tur = w * mut_hi;
  }
}

// One face: boundary overrides of the states (boundary_state_reconstruction.f90:124-131), inviscid flux times area
// (scheme.f90:68-109), viscous flux, and the face terms of the time step.  X = cell on the high side, f = its index
// along direction d (the face index), m = node count along d.
template <int NV, int SCHEME, bool VISC>
__device__ __forceinline__ void face_eval(const Params& P, const KArgs& a, int d, long long X, long long s, int f, int m, double (&L)[NV],
                                          double (&R)[NV], bool flux_on, bool need_dt, double (&F)[NV], double& lam, double& vis, double& tur) {
  const long long fs = P.L.fs;
  const double* __restrict__ q = a.q;
  if (P.interpolant != F3D_INTERP_NONE) {
    if (f == 1 && P.phys[2 * d]) {
      const bool far = P.farlike[2 * d] != 0;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const double g = q[v * fs + X - s], in = q[v * fs + X];
        if (far) { L[v] = g; R[v] = g; } else { L[v] = 0.5 * (g + in); }
      }
    }
    if (f == m && P.phys[2 * d + 1]) {
      const bool far = P.farlike[2 * d + 1] != 0;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const double in = q[v * fs + X - s], g = q[v * fs + X];
        if (far) { L[v] = g; R[v] = g; } else { R[v] = 0.5 * (in + g); }
      }
    }
  }
  const double* __restrict__ gA = a.geom + (long long)(G_IA + 4 * d) * fs;
  const double A = gA[X], nx = gA[fs + X], ny = gA[2 * fs + X], nz = gA[3 * fs + X];
  const double mask = (f == 1) ? P.zlo[d] : ((f == m) ? P.zhi[d] : 1.0);
  const double cbar = inviscid_flux<NV>(SCHEME >= 0 ? SCHEME : P.scheme, P.gm, P.MInf, L, R, A, nx, ny, nz, mask, flux_on, need_dt, F);
  if (need_dt) {   // time.f90:159-237: both cells of a face use the velocity of the cell on its high side
    const double vn = fabs((q[1 * fs + X] * nx) + (q[2 * fs + X] * ny) + (q[3 * fs + X] * nz));
    lam = A * (vn + cbar);
  }
  if (VISC) viscous_face<NV>(P, a, X - s, X, A, nx, ny, nz, (NV == 7) && flux_on, need_dt, F, vis, tur);
}

template <int NV, int INTERP, int SCHEME, bool VISC>
__global__ void __launch_bounds__(NT, 1) k_sweep(const Params P, const KArgs a) {
  constexpr bool SST = (NV == 7);
  constexpr int NF = NV + 3;
  extern __shared__ double smem[];
  double* const sm_hi = smem;                          // [NV][SLOT_HF]
  double* const sm_lo = sm_hi + NV * SLOT_HF;          // [NV][SLOT_LO]
  double* const sm_F = sm_lo + NV * SLOT_LO;           // [NF][SLOT_HF]
  const Layout& Ly = P.L;
  const long long fs = Ly.fs;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int i0 = 1 + blockIdx.x * TX, j0 = 1 + blockIdx.y * TY;
  const int kb = 1 + blockIdx.z * a.kchunk, ke = min(kb + a.kchunk, Ly.kmx);   // planes kb .. ke-1
  const bool need_dt = a.first_stage != 0;
  const bool flux_on_k = Ly.kmx != 2;   // H = 0 when kmx == 2 (ausm.f90:205-210)
  const bool k_active = flux_on_k || VISC || need_dt;

  // ---- role of this thread ------------------------------------------------------------------------------------------
  int i, j;
  bool own = false;                 // finishes a cell in phase 3
  bool rec[2] = {false, false};     // reconstructs along i / j
  bool fac[2] = {false, false};     // evaluates its low face along i / j
  int whi[2] = {0, 0}, rhi[2] = {0, 0}, wf[2] = {0, 0}, plo[2] = {0, 0};   // smem slots (relative to the direction's base)
  if (wid < TY) {
    const int tx = lane, ty = wid;
    i = i0 + tx; j = j0 + ty;
    own = (i <= Ly.imx - 1) && (j <= Ly.jmx - 1);
    rec[0] = fac[0] = (j <= Ly.jmx - 1) && (i <= Ly.imx);
    rec[1] = fac[1] = (i <= Ly.imx - 1) && (j <= Ly.jmx);
    whi[0] = ty * (TX + 1) + tx + 1; rhi[0] = ty * (TX + 1) + tx; wf[0] = ty * (TX + 1) + tx; plo[0] = tid;
    whi[1] = (ty + 1) * TX + tx; rhi[1] = ty * TX + tx; wf[1] = ty * TX + tx; plo[1] = NMAIN + tid;
  } else if (wid == TY) {           // the two i columns next to the tile: lanes 0..TY-1 low side, TY..2TY-1 high side
    const int r = lane % TY, side = lane / TY;
    i = (side == 0) ? i0 - 1 : i0 + TX; j = j0 + r;
    const bool act = (side < 2) && (j <= Ly.jmx - 1) && (i <= Ly.imx);
    rec[0] = act; fac[0] = act && side == 1;
    whi[0] = r * (TX + 1) + 0;      // only the low side's hi value is read by anybody (slot 0 of the row)
    rhi[0] = r * (TX + 1) + TX; wf[0] = r * (TX + 1) + TX; plo[0] = 3 * NMAIN + r;
    if (i > Ly.imx) i = Ly.imx;
    if (j > Ly.jmx) j = Ly.jmx;
  } else {                          // low (wid == TY+1) and high (wid == TY+2) j rows next to the tile
    const bool high = wid == TY + 2;
    i = i0 + lane; j = high ? j0 + TY : j0 - 1;
    const bool act = (i <= Ly.imx - 1) && (j <= Ly.jmx);
    rec[1] = act; fac[1] = act && high;
    whi[1] = lane;                  // low row: slot row 0
    rhi[1] = TY * TX + lane; wf[1] = TY * TX + lane; plo[1] = 3 * NMAIN + 32 + lane;
    if (i > Ly.imx) i = Ly.imx;
    if (j > Ly.jmx) j = Ly.jmx;
  }
  if (wid < TY) { if (i > Ly.imx) i = Ly.imx; if (j > Ly.jmx) j = Ly.jmx; }
  const bool hi_side_halo = (wid == TY && (lane / TY) == 1) || (wid == TY + 2);   // its own hi value is never read

  const double* __restrict__ q = a.q;
  const double* __restrict__ vol = a.geom + (long long)G_VOL * fs;
  double nrm[NV + 1];
#pragma unroll
  for (int v = 0; v <= NV; ++v) nrm[v] = 0.0;

  for (int k = kb - 2; k < ke; ++k) {
    const int par = (k - kb) & 1;                   // ping-pong slot of the k direction
    const bool inplane = k >= kb;
    const long long c = Ly.idx(i, j, k);
    // ---- phase 1: reconstruction --------------------------------------------------------------------------------------
#pragma unroll 1
    for (int d = 0; d < 3; ++d) {
      bool doit, need_lo = true;
      long long cr, s;
      int pos, mx, slot_hi, slot_lo;
      if (d == 0) { doit = inplane && rec[0]; need_lo = fac[0]; cr = c; s = 1; pos = i; mx = Ly.imx; slot_hi = whi[0]; slot_lo = plo[0]; }
      else if (d == 1) { doit = inplane && rec[1]; need_lo = fac[1]; cr = c; s = Ly.sj; pos = j; mx = Ly.jmx; slot_hi = SLOT_I + whi[1]; slot_lo = plo[1]; }
      else { doit = own && k_active; cr = c + Ly.sk; s = Ly.sk; pos = k + 1; mx = Ly.kmx; slot_hi = SLOT_I + SLOT_J + par * NMAIN + tid; slot_lo = 2 * NMAIN + tid; }
      if (!doit) continue;
      double hi[NV], lo[NV];
      line_cell_values<NV, INTERP>(P, q, vol, cr, s, pos, mx, d, hi, lo);
      if (!(hi_side_halo && d < 2)) {
#pragma unroll
        for (int v = 0; v < NV; ++v) sm_hi[v * SLOT_HF + slot_hi] = hi[v];
      }
      if (need_lo) {
#pragma unroll
        for (int v = 0; v < NV; ++v) sm_lo[v * SLOT_LO + slot_lo] = lo[v];
      }
    }
    __syncthreads();
    // ---- phase 2: faces --------------------------------------------------------------------------------------------------
#pragma unroll 1
    for (int d = 0; d < 3; ++d) {
      bool doit, flux_on = true;
      long long X, s;
      int f, m, slot_L, slot_R, slot_F;
      if (d == 0) { doit = inplane && fac[0]; X = c; s = 1; f = i; m = Ly.imx; slot_L = rhi[0]; slot_R = plo[0]; slot_F = wf[0]; }
      else if (d == 1) { doit = inplane && fac[1]; X = c; s = Ly.sj; f = j; m = Ly.jmx; slot_L = SLOT_I + rhi[1]; slot_R = plo[1]; slot_F = SLOT_I + wf[1]; }
      else {
        doit = own && k_active && k >= kb - 1; X = c + Ly.sk; s = Ly.sk; f = k + 1; m = Ly.kmx; flux_on = flux_on_k;
        slot_L = SLOT_I + SLOT_J + (par ^ 1) * NMAIN + tid; slot_R = 2 * NMAIN + tid; slot_F = SLOT_I + SLOT_J + par * NMAIN + tid;
      }
      if (!doit) continue;
      double L[NV], R[NV], F[NV], lam = 0.0, vis = 0.0, tur = 0.0;
#pragma unroll
      for (int v = 0; v < NV; ++v) { L[v] = sm_hi[v * SLOT_HF + slot_L]; R[v] = sm_lo[v * SLOT_LO + slot_R]; }
      face_eval<NV, SCHEME, VISC>(P, a, d, X, s, f, m, L, R, flux_on, need_dt, F, lam, vis, tur);
#pragma unroll
      for (int v = 0; v < NV; ++v) sm_F[v * SLOT_HF + slot_F] = F[v];
      if (need_dt) {
        sm_F[NV * SLOT_HF + slot_F] = lam;
        if (VISC) sm_F[(NV + 1) * SLOT_HF + slot_F] = vis;
        if (VISC && SST) sm_F[(NV + 2) * SLOT_HF + slot_F] = tur;
      }
    }
    __syncthreads();
    // ---- phase 3: the cell ---------------------------------------------------------------------------------------------------
    if (!(inplane && own)) continue;
    const int tx = lane, ty = wid;
    const int sl[3] = {ty * (TX + 1) + tx, SLOT_I + ty * TX + tx, SLOT_I + SLOT_J + (par ^ 1) * NMAIN + tid};        // low faces
    const int sh[3] = {ty * (TX + 1) + tx + 1, SLOT_I + (ty + 1) * TX + tx, SLOT_I + SLOT_J + par * NMAIN + tid};    // high faces
    double res[NV];
    double merr = 0.0;
#pragma unroll
    for (int v = 0; v < NV; ++v) res[v] = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (d == 2 && !k_active) continue;
      const int p = (d == 0) ? i : (d == 1 ? j : k);
      const int m = (d == 0) ? Ly.imx : (d == 1 ? Ly.jmx : Ly.kmx);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const double Fl = sm_F[v * SLOT_HF + sl[d]], Fh = sm_F[v * SLOT_HF + sh[d]];
        res[v] = res[v] + (Fh - Fl);   // scheme.f90:133-135
        if (v == 0) {                  // resnorm.f90:190-198
          if (p == 1) merr += Fl;
          if (p == m - 1) merr -= Fh;
        }
      }
    }
    {
      bool bad = false;
#pragma unroll
      for (int v = 0; v < NV; ++v) bad |= isnan(res[v]);
      if (bad) flag_error(a.err, F3D_ERR_NAN_FLUX, i, j, k);
    }
    const double volc = vol[c];
    double qc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) qc[v] = q[v * fs + c];
    if (SST && VISC) {   // source.f90:214-268
      double g[6][3];
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) {
        if (cc == 3) continue;
        g[cc][0] = a.grad[(3 * cc + 0) * fs + c]; g[cc][1] = a.grad[(3 * cc + 1) * fs + c]; g[cc][2] = a.grad[(3 * cc + 2) * fs + c];
      }
      const double mut = a.mu[fs + c];
      const double density = qc[0], tk = qc[5], tw = qc[6];
      const double wx = g[2][1] - g[1][2], wy = g[0][2] - g[2][0], wz = g[1][0] - g[0][1];
      const double vort = sqrt(wx * wx + wy * wy + wz * wz);
      double CD = 2 * density * kSigmaW2 * (g[4][0] * g[5][0] + g[4][1] * g[5][1] + g[4][2] * g[5][2]) * rcp64(tw);
      CD = dmax(CD, P.cd_floor);
      const double F1 = a.mu[2 * fs + c];
      const double gama = P.gama1 * F1 + P.gama2 * (1. - F1);
      const double beta = kBeta1 * F1 + kBeta2 * (1. - F1);
      const double D_k = kBstar * density * tw * tk;
      const double D_w = beta * density * (tw * tw);
      const double divergence = g[0][0] + g[1][1] + g[2][2];
      double P_k = mut * (vort * vort) - ((2.0 / 3.0) * density * tk * divergence);
      P_k = dmin(P_k, P.pk_limiter * D_k);
      const double P_w = (density * gama * rcp64(mut)) * P_k;
      const double lamda = (1. - F1) * CD;
      const double S_k = (P_k - D_k) * volc;
      const double S_w = (P_w - D_w + lamda) * volc;
      res[5] = res[5] - S_k;
      res[6] = res[6] - S_w;
    }

    double dtc = 0.0;
    if (need_dt) {
      if (P.time_stepping == 1 && P.global_time_step > 0) {
        dtc = P.global_time_step;
      } else {
        const double* lamv = sm_F + NV * SLOT_HF;
        const double lmxsum = lamv[sl[0]] + lamv[sl[1]] + lamv[sl[2]] + lamv[sh[0]] + lamv[sh[1]] + lamv[sh[2]];
        dtc = rcp64(lmxsum);
        dtc = dtc * volc * P.CFL;
        if (VISC) {
          const double* visv = sm_F + (NV + 1) * SLOT_HF;
          double s = visv[sl[0]] + visv[sl[1]] + visv[sl[2]] + visv[sh[0]] + visv[sh[1]] + visv[sh[2]];
          s = P.gm * s * P.inv_Pr;
          s = 2. * rcp64(s + (2. * P.CFL * volc * rcp64(dtc)));
          dtc = P.CFL * (s * volc);
          if (SST) {
            const double* turv = sm_F + (NV + 2) * SLOT_HF;
            double t = turv[sl[0]] + turv[sl[1]] + turv[sl[2]] + turv[sh[0]] + turv[sh[1]] + turv[sh[2]];
            t = P.gm * t * P.inv_tPr;
            t = 2. * rcp64(t + (2. * P.CFL * volc * rcp64(dtc)));
            dtc = P.CFL * (t * volc);
          }
        }
      }
      a.dt[c] = dtc;
    } else if (a.mode == MODE_UPDATE) {
      dtc = a.dt[c];
    }

    if (a.mode == MODE_RESIDUE_ONLY) {
#pragma unroll
      for (int v = 0; v < NV; ++v) a.residue[v * fs + c] = res[v];
    } else {   // update.f90:371-485
      double u1[NV], R[NV], u2[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) R[v] = res[v];
      u1[0] = a.quse[c];
#pragma unroll
      for (int v = 1; v < NV; ++v) u1[v] = a.quse[v * fs + c] * u1[0];
      u1[4] = (u1[4] * P.inv_gm1 + 0.5 * (u1[1] * u1[1] + u1[2] * u1[2] + u1[3] * u1[3])) * rcp64(u1[0]) + 0.;
      if (SST) {
        const double F1 = a.mu[2 * fs + c];
        const double beta = kBeta1 * F1 + (1. - F1) * kBeta2;
        R[5] = R[5] * rcp64(1 + (beta * qc[6] * dtc));
        R[6] = R[6] * rcp64(1 + (2 * beta * qc[6] * dtc));
      }
      if (a.have_store) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const double rn = a.rstore[v * fs + c] + a.SF * R[v];
          a.rstore[v * fs + c] = rn;
          if (a.use_store_sum) R[v] = rn;
        }
      }
      const double fac_ = (a.TF * dtc * rcp64(volc));
#pragma unroll
      for (int v = 0; v < NV; ++v) u2[v] = u1[v] - R[v] * fac_;
      const double iu = 1.0 / u2[0];   // IEEE: u2[0] may be <= 0 or NaN here and must reach the check below unchanged
#pragma unroll
      for (int v = 1; v < NV; ++v) u2[v] = u2[v] * iu;
      u2[4] = (P.gm - 1.) * u2[0] * (u2[4] - (0.5 * (u2[1] * u2[1] + u2[2] * u2[2] + u2[3] * u2[3])) - 0.);
      bool bad = (u2[0] < 0.) || (u2[4] < 0.);
#pragma unroll
      for (int v = 0; v < NV; ++v) bad |= isnan(u2[v]);
      if (bad) {
        flag_error(a.err, F3D_ERR_NEGATIVE_STATE, i, j, k);
#pragma unroll
        for (int v = 0; v < NV; ++v) a.qnew[v * fs + c] = qc[v];
      } else {
#pragma unroll
        for (int v = 0; v < 5; ++v) a.qnew[v * fs + c] = u2[v];
        if (SST) {
          a.qnew[5 * fs + c] = (u2[5] >= 0.) ? u2[5] : qc[5];
          a.qnew[6 * fs + c] = (u2[6] >= 0.) ? u2[6] : qc[6];
        }
      }
    }
    if (a.want_norms) {   // resnorm.f90:187-198
      nrm[0] += merr;
#pragma unroll
      for (int v = 0; v < NV; ++v) nrm[v + 1] += res[v] * res[v];
    }
  }

  if (a.want_norms) {   // per-CTA partial: warp shuffle, then the block
    __syncthreads();
    double* sred = smem;   // [NV+1][NT/32]
#pragma unroll
    for (int v = 0; v <= NV; ++v) {
      double x = nrm[v];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      if (lane == 0) sred[v * (NT / 32) + wid] = x;
    }
    __syncthreads();
    if (tid <= NV) {
      double x = 0.0;
      for (int w = 0; w < TY; ++w) x += sred[tid * (NT / 32) + w];
      const long long cta = blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z);
      a.red[cta * (NV + 1) + tid] = x;
    }
  }
}

// final reduction of the per-CTA partials in a fixed order (deterministic), scaled like get_absolute_resnorm
__global__ void k_norm_final(const double* __restrict__ red, int n_cta, int nvp1, const double* scale /* nvp1 */, double* out) {
  __shared__ double sm[32];
  const int v = blockIdx.x;
  double x = 0.0;
  for (int b = threadIdx.x; b < n_cta; b += blockDim.x) x += red[(long long)b * nvp1 + v];
  for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
    out[v] = (v == 0) ? (t / scale[0]) : (t / (scale[v] * scale[v]));
  }
}

// k planes per CTA: long enough to amortise the two priming iterations, short enough for >= ~4 waves of CTAs
static int pick_kchunk(const Layout& L) {
  const int nk = L.kmx - 1;
  const long long tiles = (long long)((L.imx - 1 + TX - 1) / TX) * ((L.jmx - 1 + TY - 1) / TY);
  int chunk = nk;
  while (chunk > 16 && tiles * ((nk + chunk - 1) / chunk) < 148 * 4) chunk = (chunk + 1) / 2;
  return chunk;
}

int residual_grid_ctas(const Layout& L) {
  const int chunk = pick_kchunk(L);
  return ((L.imx - 1 + TX - 1) / TX) * ((L.jmx - 1 + TY - 1) / TY) * ((L.kmx - 1 + chunk - 1) / chunk);
}

template <int NV, int INTERP, int SCHEME, bool VISC>
static int launch_one(Ctx* ctx, KArgs& a) {
  const Layout& L = ctx->P.L;
  a.kchunk = pick_kchunk(L);
  dim3 grid((L.imx - 1 + TX - 1) / TX, (L.jmx - 1 + TY - 1) / TY, (L.kmx - 1 + a.kchunk - 1) / a.kchunk);
  const size_t shm = sizeof(double) * sweep_smem_doubles(NV);
  static bool attr_set[64] = {false};   // per instantiation and device
  if (!attr_set[ctx->device & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k_sweep<NV, INTERP, SCHEME, VISC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    if (e != cudaSuccess) return F3D_ERR_CUDA;
    attr_set[ctx->device & 63] = true;
  }
  k_sweep<NV, INTERP, SCHEME, VISC><<<grid, NT, shm, ctx->stream>>>(ctx->P, a);
  ctx->launches++;
  return 0;
}

template <int NV, bool VISC>
static int launch_interp(Ctx* ctx, KArgs& a) {
  switch (ctx->P.interpolant) {
    case F3D_INTERP_NONE: return launch_one<NV, F3D_INTERP_NONE, -1, VISC>(ctx, a);
    case F3D_MUSCL:
      if (ctx->P.scheme == F3D_AUSM) return launch_one<NV, F3D_MUSCL, F3D_AUSM, VISC>(ctx, a);   // the headline configuration
      return launch_one<NV, F3D_MUSCL, -1, VISC>(ctx, a);
    case F3D_PPM: return launch_one<NV, F3D_PPM, -1, VISC>(ctx, a);
    case F3D_WENO: return launch_one<NV, F3D_WENO, -1, VISC>(ctx, a);
    case F3D_WENO_NM: return launch_one<NV, F3D_WENO_NM, -1, VISC>(ctx, a);
  }
  return F3D_ERR_UNSUPPORTED;
}

int launch_residual(Ctx* ctx, int mode, double TF, double SF, int use_store_sum, int first_stage, int want_norms) {
  KArgs a{};
  a.q = ctx->qp;
  const bool have_store = ctx->cfg.time_accuracy == F3D_T_RK2 || ctx->cfg.time_accuracy == F3D_T_RK4;
  a.quse = (mode == MODE_UPDATE && have_store) ? ctx->ustore : ctx->qp;
  a.qnew = ctx->qp2;
  a.residue = ctx->residue;
  a.rstore = (mode == MODE_UPDATE && have_store) ? ctx->rstore : nullptr;
  a.dt = ctx->dt;
  a.geom = ctx->geom;
  a.grad = ctx->grad;
  a.mu = ctx->mu;
  a.red = ctx->red;
  a.err = ctx->err_dev;
  a.mode = mode; a.first_stage = first_stage; a.want_norms = want_norms;
  a.have_store = (mode == MODE_UPDATE && have_store) ? 1 : 0;
  a.use_store_sum = use_store_sum;
  a.TF = TF; a.SF = SF;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->timing) {
    if (ctx->ev_used == ctx->ev_pool.size()) {
      cudaEvent_t x, y; cudaEventCreate(&x); cudaEventCreate(&y);
      ctx->ev_pool.emplace_back(x, y);
    }
    e0 = ctx->ev_pool[ctx->ev_used].first; e1 = ctx->ev_pool[ctx->ev_used].second; ctx->ev_used++;
    cudaEventRecord(e0, ctx->stream);
  }
  int rc;
  if (ctx->P.viscous) rc = ctx->P.sst ? launch_interp<7, true>(ctx, a) : launch_interp<5, true>(ctx, a);
  else rc = launch_interp<5, false>(ctx, a);
  if (ctx->timing) cudaEventRecord(e1, ctx->stream);
  if (rc) return rc;
  F3D_CUDA(cudaGetLastError());
  return 0;
}

int launch_norms(Ctx* ctx, int slot) {
  const int nvp1 = ctx->P.L.nv + 1;
  k_norm_final<<<nvp1, 256, 0, ctx->stream>>>(ctx->red, ctx->red_blocks, nvp1, ctx->norms_dev + 1024, ctx->norms_dev + (long long)slot * nvp1);
  ctx->launches++;
  F3D_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace f3d
