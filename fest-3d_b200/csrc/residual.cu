// Fused residual kernel, generation 0 ("cell gather"): one thread per interior cell gathers its stencil, reconstructs
// the face states of its six faces, evaluates inviscid (+ viscous) fluxes, forms the residual, adds the SST source and
// -- in update mode -- the local time step, the point-implicit turbulence scaling, the RK accumulate and the
// conservative update, writing the new primitive state to the second state buffer.  No face-state, flux or residual
// array reaches HBM on the update path; the residual-norm partials are reduced in-kernel (warp shuffle + block).
//
// Reference pipeline reproduced per cell (src/update.f90:534-545, 228-491; src/time.f90:122-246,366-531;
// src/face/flux/convective/scheme.f90:111-141; src/source.f90:158-270; src/resnorm.f90:171-199).
#include "ctx.hpp"
#include "physics.cuh"

namespace f3d {

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
  double* __restrict__ red;           // per-CTA partials [(nv+1) * n_cta] or nullptr
  int* err;
  int mode, first_stage, want_norms, have_store, use_store_sum;
  double TF, SF;
};

// Values a cell contributes to its two faces along one direction, all variables.  `pos` is the cell's index along
// the direction; the first / last interior cell next to a physical boundary is re-done with the boundary formula
// when ppm_flag is set (boundary_state_reconstruction.f90:93-123).
template <int NV, int INTERP>
__device__ __forceinline__ void line_cell_values(const Params& P, const double* __restrict__ q, const double* __restrict__ vol,
                                                 long long c, long long s, int pos, int mx, int dir, bool phys_lo, bool phys_hi,
                                                 double (&to_hi)[NV], double (&to_lo)[NV]) {
  const bool redo = (INTERP != F3D_INTERP_NONE) && P.ppm_flag && ((pos == 1 && phys_lo) || (pos == mx - 1 && phys_hi));
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

// everything the viscous face flux needs from one cell
template <int NV>
struct CellVisc {
  double q[NV];
  double g[(NV == 7) ? 6 : 4][3];
  double mu, mut, F1, cx, cy, cz;
};

template <int NV>
__device__ __forceinline__ void load_cell_visc(const Params& P, const KArgs& a, long long c, CellVisc<NV>& o) {
  constexpr int NG = (NV == 7) ? 6 : 4;
  const long long fs = P.L.fs;
#pragma unroll
  for (int v = 0; v < NV; ++v) o.q[v] = a.q[v * fs + c];
#pragma unroll
  for (int cc = 0; cc < NG; ++cc) {
    o.g[cc][0] = a.grad[(3 * cc + 0) * fs + c];
    o.g[cc][1] = a.grad[(3 * cc + 1) * fs + c];
    o.g[cc][2] = a.grad[(3 * cc + 2) * fs + c];
  }
  o.mu = a.mu[c];
  o.mut = (NV == 7) ? a.mu[fs + c] : 0.0;
  o.F1 = (NV == 7) ? a.mu[2 * fs + c] : 0.0;
  o.cx = a.geom[(long long)G_CX * fs + c]; o.cy = a.geom[(long long)G_CY * fs + c]; o.cz = a.geom[(long long)G_CZ * fs + c];
}

// F <- (F - laminar) - sst  for one face between cells lo and hi (viscous.f90:316-323, 437-442)
template <int NV>
__device__ __forceinline__ void apply_viscous(const Params& P, const CellVisc<NV>& lo, const CellVisc<NV>& hi,
                                              double A, double nx, double ny, double nz, bool sst_on, double (&F)[NV]) {
  constexpr bool SST = (NV == 7);
  constexpr int NG = SST ? 6 : 4;
  const double dx = hi.cx - lo.cx, dy = hi.cy - lo.cy, dz = hi.cz - lo.cz;
  const double d_LR = sqrt(dx * dx + dy * dy + dz * dz);
  double del[NG];
  del[0] = hi.q[1] - lo.q[1]; del[1] = hi.q[2] - lo.q[2]; del[2] = hi.q[3] - lo.q[3];
  {
    const double T_LE = lo.q[4] / (lo.q[0] * P.R_gas), T_RE = hi.q[4] / (hi.q[0] * P.R_gas);
    del[3] = T_RE - T_LE;
  }
  if (SST) { del[4] = hi.q[5] - lo.q[5]; del[5] = hi.q[6] - lo.q[6]; }
  double G[NG][3];
#pragma unroll
  for (int c = 0; c < NG; ++c) {
    const double ax = 0.5 * (lo.g[c][0] + hi.g[c][0]), ay = 0.5 * (lo.g[c][1] + hi.g[c][1]), az = 0.5 * (lo.g[c][2] + hi.g[c][2]);
    const double nc = (del[c] - (ax * dx + ay * dy + az * dz)) / d_LR;
    G[c][0] = ax + (nc * dx / d_LR);
    G[c][1] = ay + (nc * dy / d_LR);
    G[c][2] = az + (nc * dz / d_LR);
  }
  const double mu_f = 0.5 * (lo.mu + hi.mu);
  const double mut_f = SST ? 0.5 * (lo.mut + hi.mut) : 0.0;
  const double tmu = mu_f + mut_f;
  const double div3 = (G[0][0] + G[1][1] + G[2][2]) / 3.;
  const double Txx = 2. * tmu * (G[0][0] - div3), Tyy = 2. * tmu * (G[1][1] - div3), Tzz = 2. * tmu * (G[2][2] - div3);
  const double Txy = tmu * (G[1][0] + G[0][1]), Txz = tmu * (G[2][0] + G[0][2]), Tyz = tmu * (G[2][1] + G[1][2]);
  const double Kh = (mu_f / P.Pr + mut_f / P.tPr) * P.gm * P.R_gas / (P.gm - 1);
  const double Qx = Kh * G[3][0], Qy = Kh * G[3][1], Qz = Kh * G[3][2];
  const double uf = 0.5 * (lo.q[1] + hi.q[1]), vf = 0.5 * (lo.q[2] + hi.q[2]), wf = 0.5 * (lo.q[3] + hi.q[3]);
  F[1] = F[1] - ((Txx * nx + Txy * ny + Txz * nz) * A);
  F[2] = F[2] - ((Txy * nx + Tyy * ny + Tyz * nz) * A);
  F[3] = F[3] - ((Txz * nx + Tyz * ny + Tzz * nz) * A);
  F[4] = F[4] - (A * (((Txx * uf + Txy * vf + Txz * wf + Qx) * nx) + ((Txy * uf + Tyy * vf + Tyz * wf + Qy) * ny) +
                      ((Txz * uf + Tyz * vf + Tzz * wf + Qz) * nz)));
  if (SST && sst_on) {
    const double F1 = 0.5 * (lo.F1 + hi.F1);
    const double sk = kSigmaK1 * F1 + kSigmaK2 * (1.0 - F1);
    const double sw = kSigmaW1 * F1 + kSigmaW2 * (1.0 - F1);
    const double rhof = 0.5 * (lo.q[0] + hi.q[0]);
    const double tkf = 0.5 * (lo.q[NV - 2] + hi.q[NV - 2]);
    const double Tk = -2.0 * rhof * tkf / 3.0;
    const double dk = (A * ((mu_f + sk * mut_f) * (G[NG - 2][0] * nx + G[NG - 2][1] * ny + G[NG - 2][2] * nz)));
    const double dw = (A * ((mu_f + sw * mut_f) * (G[NG - 1][0] * nx + G[NG - 1][1] * ny + G[NG - 1][2] * nz)));
    F[1] = F[1] - (Tk * nx * A);
    F[2] = F[2] - (Tk * ny * A);
    F[3] = F[3] - (Tk * nz * A);
    F[4] = F[4] - dk;
    F[NV - 2] = F[NV - 2] - dk;
    F[NV - 1] = F[NV - 1] - dw;
  }
}

template <int NV, int INTERP, bool VISC>
__global__ void __launch_bounds__(128) k_residual(const Params P, const KArgs a) {
  constexpr bool SST = (NV == 7);
  const Layout& Ly = P.L;
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  const int k = 1 + blockIdx.z;
  const bool active = (i <= Ly.imx - 1) && (j <= Ly.jmx - 1);
  double res[NV];
  double merr = 0.0;
#pragma unroll
  for (int v = 0; v < NV; ++v) res[v] = 0.0;

  if (active) {
    const long long c = Ly.idx(i, j, k);
    const long long fs = Ly.fs;
    const double* __restrict__ q = a.q;
    const double* __restrict__ vol = a.geom + (long long)G_VOL * fs;
    const bool need_dt = a.first_stage != 0;
    double lam_lo[3] = {0, 0, 0}, lam_hi[3] = {0, 0, 0};   // A*(|V.n| + c)
    double vis_lo[3] = {0, 0, 0}, vis_hi[3] = {0, 0, 0};   // A*mu/(rho |dr.n|)
    double tur_lo[3] = {0, 0, 0}, tur_hi[3] = {0, 0, 0};   // A*mu_t/(rho |dr.n|)
    CellVisc<NV> cv;
    if (VISC) load_cell_visc<NV>(P, a, c, cv);

#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const bool flux_on = !(d == 2 && Ly.kmx == 2);   // H = 0 when kmx == 2 (ausm.f90:205-210)
      if (!flux_on && !VISC && !need_dt) continue;
      const long long s = (d == 0) ? 1 : (d == 1 ? Ly.sj : Ly.sk);
      const int p = (d == 0) ? i : (d == 1 ? j : k);
      const int m = (d == 0) ? Ly.imx : (d == 1 ? Ly.jmx : Ly.kmx);
      const bool phys_lo = P.phys[2 * d] != 0, phys_hi = P.phys[2 * d + 1] != 0;
      double Ll[NV], Rl[NV], Lh[NV], Rh[NV];
      {
        double tmp[NV];
        line_cell_values<NV, INTERP>(P, q, vol, c - s, s, p - 1, m, d, phys_lo, phys_hi, Ll, tmp);
        line_cell_values<NV, INTERP>(P, q, vol, c, s, p, m, d, phys_lo, phys_hi, Lh, Rl);
        line_cell_values<NV, INTERP>(P, q, vol, c + s, s, p + 1, m, d, phys_lo, phys_hi, tmp, Rh);
      }
      if (INTERP != F3D_INTERP_NONE) {   // boundary_state_reconstruction.f90:124-131
        if (p == 1 && phys_lo) {
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const double g = q[v * fs + c - s], in = q[v * fs + c];
            if (P.farlike[2 * d]) { Ll[v] = g; Rl[v] = g; } else { Ll[v] = 0.5 * (g + in); }
          }
        }
        if (p == m - 1 && phys_hi) {
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const double g = q[v * fs + c + s], in = q[v * fs + c];
            if (P.farlike[2 * d + 1]) { Lh[v] = g; Rh[v] = g; } else { Rh[v] = 0.5 * (in + g); }
          }
        }
      }
      const double* __restrict__ gA = a.geom + (long long)(G_IA + 4 * d) * fs;
      const double Al = gA[c], nxl = gA[fs + c], nyl = gA[2 * fs + c], nzl = gA[3 * fs + c];
      const double Ah = gA[c + s], nxh = gA[fs + c + s], nyh = gA[2 * fs + c + s], nzh = gA[3 * fs + c + s];
      double Fl[NV], Fh[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) { Fl[v] = 0.0; Fh[v] = 0.0; }
      if (flux_on) {
        const double ml = (p == 1) ? P.zlo[d] : 1.0;
        const double mh = (p == m - 1) ? P.zhi[d] : 1.0;
        inviscid_flux<NV>(P.scheme, P.gm, P.MInf, Ll, Rl, Al, nxl, nyl, nzl, ml, Fl);
        inviscid_flux<NV>(P.scheme, P.gm, P.MInf, Lh, Rh, Ah, nxh, nyh, nzh, mh, Fh);
      }
      if (need_dt) {   // time.f90:159-237
        const double vn_lo = fabs((q[1 * fs + c] * nxl) + (q[2 * fs + c] * nyl) + (q[3 * fs + c] * nzl));
        const double vn_hi = fabs((q[1 * fs + c + s] * nxh) + (q[2 * fs + c + s] * nyh) + (q[3 * fs + c + s] * nzh));
        lam_lo[d] = Al * (vn_lo + face_sound_speed<NV>(P.gm, Ll, Rl));
        lam_hi[d] = Ah * (vn_hi + face_sound_speed<NV>(P.gm, Lh, Rh));
      }
      if (VISC) {
        CellVisc<NV> nb;
        const bool sst_on = SST && flux_on;   // SST K flux skipped when kmx == 2 (viscous.f90:101-105)
        load_cell_visc<NV>(P, a, c - s, nb);
        apply_viscous<NV>(P, nb, cv, Al, nxl, nyl, nzl, sst_on, Fl);
        if (need_dt) {   // time.f90:396-407: low faces use the cell's own mu and density
          const double dn = fabs(((nb.cx - cv.cx) * nxl) + ((nb.cy - cv.cy) * nyl) + ((nb.cz - cv.cz) * nzl));
          vis_lo[d] = Al * (cv.mu / (cv.q[0] * dn));
          if (SST) tur_lo[d] = Al * (cv.mut / (cv.q[0] * dn));
        }
        load_cell_visc<NV>(P, a, c + s, nb);
        apply_viscous<NV>(P, cv, nb, Ah, nxh, nyh, nzh, sst_on, Fh);
        if (need_dt) {   // time.f90:410-421: high faces use the neighbour's mu and density
          const double dn = fabs(((cv.cx - nb.cx) * nxh) + ((cv.cy - nb.cy) * nyh) + ((cv.cz - nb.cz) * nzh));
          vis_hi[d] = Ah * (nb.mu / (nb.q[0] * dn));
          if (SST) tur_hi[d] = Ah * (nb.mut / (nb.q[0] * dn));
        }
      }
      {
        bool bad = false;
#pragma unroll
        for (int v = 0; v < NV; ++v) bad |= isnan(Fl[v]) || isnan(Fh[v]);
        if (bad) flag_error(a.err, F3D_ERR_NAN_FLUX, i, j, k);
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) res[v] = res[v] + (Fh[v] - Fl[v]);   // scheme.f90:133-135
      if (p == 1) merr += Fl[0];          // resnorm.f90:190-198
      if (p == m - 1) merr -= Fh[0];
    }

    const double volc = vol[c];
    if (SST && VISC) {   // source.f90:214-268
      const double density = cv.q[0], tk = cv.q[5], tw = cv.q[6];
      const double wx = cv.g[2][1] - cv.g[1][2], wy = cv.g[0][2] - cv.g[2][0], wz = cv.g[1][0] - cv.g[0][1];
      const double vort = sqrt(wx * wx + wy * wy + wz * wz);
      double CD = 2 * density * kSigmaW2 * (cv.g[4][0] * cv.g[5][0] + cv.g[4][1] * cv.g[5][1] + cv.g[4][2] * cv.g[5][2]) / tw;
      CD = fmax(CD, P.cd_floor);
      const double F1 = cv.F1;
      const double gama = P.gama1 * F1 + P.gama2 * (1. - F1);
      const double beta = kBeta1 * F1 + kBeta2 * (1. - F1);
      const double D_k = kBstar * density * tw * tk;
      const double D_w = beta * density * (tw * tw);
      const double divergence = cv.g[0][0] + cv.g[1][1] + cv.g[2][2];
      double P_k = cv.mut * (vort * vort) - ((2.0 / 3.0) * density * tk * divergence);
      P_k = fmin(P_k, P.pk_limiter * D_k);
      const double P_w = (density * gama / cv.mut) * P_k;
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
        const double lmxsum = lam_lo[0] + lam_lo[1] + lam_lo[2] + lam_hi[0] + lam_hi[1] + lam_hi[2];
        dtc = 1. / lmxsum;
        dtc = dtc * volc * P.CFL;
        if (VISC) {
          double s = vis_lo[0] + vis_lo[1] + vis_lo[2] + vis_hi[0] + vis_hi[1] + vis_hi[2];
          s = P.gm * s / P.Pr;
          s = 2. / (s + (2. * P.CFL * volc / dtc));
          dtc = P.CFL * (s * volc);
          if (SST) {
            double t = tur_lo[0] + tur_lo[1] + tur_lo[2] + tur_hi[0] + tur_hi[1] + tur_hi[2];
            t = P.gm * t / P.tPr;
            t = 2. / (t + (2. * P.CFL * volc / dtc));
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
      double u1[NV], R[NV], u2[NV], qc[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) { qc[v] = q[v * fs + c]; R[v] = res[v]; }
      u1[0] = a.quse[c];
#pragma unroll
      for (int v = 1; v < NV; ++v) u1[v] = a.quse[v * fs + c] * u1[0];
      u1[4] = (u1[4] / (P.gm - 1.) + 0.5 * (u1[1] * u1[1] + u1[2] * u1[2] + u1[3] * u1[3])) / u1[0] + 0.;
      if (SST) {
        const double F1 = a.mu[2 * fs + c];
        const double beta = kBeta1 * F1 + (1. - F1) * kBeta2;
        R[5] = R[5] / (1 + (beta * qc[6] * dtc));
        R[6] = R[6] / (1 + (2 * beta * qc[6] * dtc));
      }
      if (a.have_store) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const double rn = a.rstore[v * fs + c] + a.SF * R[v];
          a.rstore[v * fs + c] = rn;
          if (a.use_store_sum) R[v] = rn;
        }
      }
      const double fac = (a.TF * dtc / volc);
#pragma unroll
      for (int v = 0; v < NV; ++v) u2[v] = u1[v] - R[v] * fac;
#pragma unroll
      for (int v = 1; v < NV; ++v) u2[v] = u2[v] / u2[0];
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
  }

  if (a.want_norms) {   // resnorm.f90:187-198: sum of residue^2 per variable and the boundary mass-flux imbalance
    __shared__ double sm[NV + 1][4];
    double vals[NV + 1];
    vals[0] = merr;
#pragma unroll
    for (int v = 0; v < NV; ++v) vals[v + 1] = active ? res[v] * res[v] : 0.0;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int v = 0; v <= NV; ++v) {
      double x = vals[v];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      if (lane == 0) sm[v][wid] = x;
    }
    __syncthreads();
    if (tid <= NV) {
      const int nw = (blockDim.x * blockDim.y + 31) >> 5;
      double x = 0.0;
      for (int w = 0; w < nw; ++w) x += sm[tid][w];
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

template <int NV, int INTERP, bool VISC>
static int launch_one(Ctx* ctx, const KArgs& a) {
  const Layout& L = ctx->P.L;
  dim3 block(32, 4, 1);
  dim3 grid((L.imx - 1 + block.x - 1) / block.x, (L.jmx - 1 + block.y - 1) / block.y, L.kmx - 1);
  k_residual<NV, INTERP, VISC><<<grid, block, 0, ctx->stream>>>(ctx->P, a);
  ctx->launches++;
  return 0;
}

template <int NV, bool VISC>
static int launch_interp(Ctx* ctx, const KArgs& a) {
  switch (ctx->P.interpolant) {
    case F3D_INTERP_NONE: return launch_one<NV, F3D_INTERP_NONE, VISC>(ctx, a);
    case F3D_MUSCL: return launch_one<NV, F3D_MUSCL, VISC>(ctx, a);
    case F3D_PPM: return launch_one<NV, F3D_PPM, VISC>(ctx, a);
    case F3D_WENO: return launch_one<NV, F3D_WENO, VISC>(ctx, a);
    case F3D_WENO_NM: return launch_one<NV, F3D_WENO_NM, VISC>(ctx, a);
  }
  return F3D_ERR_UNSUPPORTED;
}

int residual_grid_ctas(const Layout& L) {
  return ((L.imx - 1 + 31) / 32) * ((L.jmx - 1 + 3) / 4) * (L.kmx - 1);
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
