// Fused residual / time-step / update kernel, generation 2 ("plane sweep with staged cell records").
//
// One CTA owns a TX x TY column of cells and marches through a chunk of k planes.  The per-cell record of a plane
// (primitive state, Green-Gauss gradients, mu / mu_t / F1, cell centre) is staged ONCE in shared memory -- the tile plus a
// one-cell ring, two planes deep (k and k+1), filled by cp.async while the previous plane is being worked on -- so every
// neighbour access of the reconstruction, of the viscous face flux and of the time-step terms is an LDS with an
// immediate offset instead of a dependent global load (generation 1 was stalled on exactly those: profiles/r01_g1_*).
// Per plane every cell reconstructs its face values ONCE per direction, every face flux (inviscid + viscous + the face
// terms of the local time step) is evaluated ONCE by the thread of the cell on its high side:
//
//   i, j   reconstruct -> hi value to smem | barrier | L = hi of the low neighbour, R = own lo -> flux -> smem
//   k      same thread: hi value and low-face flux of the previous plane are carried in registers
//   cell   residual = (F(i+1)-F(i)) + (G(j+1)-G(j)) + (H(k+1)-H(k)), SST source, local time step, point-implicit k/omega
//          scaling, RK accumulate, conservative update, norm partials
//
// Cells just outside the tile in i and j are served by three extra "halo" warps (the two i columns, the low j row, the
// high j row), so no face is computed twice inside a tile.  The i/j direction loop is NOT unrolled (one code instance;
// the v0 kernel had 0.58 MB of SASS and was instruction-fetch bound, profiles/r01_v0_summary.md).  No face-state, flux
// or residual array reaches HBM on the update path.
//
// Reference pipeline reproduced (src/update.f90:534-545, 228-491; src/face/state/*.f90;
// src/boundary/boundary_state_reconstruction.f90:93-131; src/face/flux/convective/*.f90 and scheme.f90:111-141;
// src/viscous.f90:144-447; src/source.f90:158-270; src/time.f90:122-246,366-531; src/resnorm.f90:171-199).
#include "sweep_common.cuh"
#include <cstdlib>
#include <cstring>

namespace f3d {

#ifdef F3D_PHASE_TIMING   // development aid: per-phase clock64 totals of one main warp per CTA (scratch/phase_timing.py)
__device__ unsigned long long g_phase[16];
#define PT_DECL unsigned long long pt_t = clock64(), pt_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define PT_MARK(n) { const unsigned long long t_ = clock64(); pt_acc[n] += t_ - pt_t; pt_t = t_; }
#define PT_FLUSH if (tid == 32) { for (int n_ = 0; n_ < 8; ++n_) atomicAdd(&g_phase[n_], pt_acc[n_]); }
#else
#define PT_DECL
#define PT_MARK(n)
#define PT_FLUSH
#endif

#ifndef F3D_TY
#define F3D_TY 5
#endif
constexpr int TX = 32, TY = F3D_TY;
constexpr int NMAIN = TX * TY;
constexpr int NT = NMAIN + 96;   // + i-halo warp, low-j-halo warp, high-j-halo warp

// staged plane: (TX+2) x (TY+2) slots, slot = (ty+1)*PW + (tx+1); the q fields carry NOUT extra slots for the second
// ring cells the halo threads' own reconstruction reads
constexpr int PW = TX + 2;
constexpr int PS = PW * (TY + 2);
constexpr int NOUT = 2 * TY + 2 * TX;
constexpr int PSQ = PS + NOUT;
// exchange buffers ([variable][slot]): i faces TY x (TX+1), j faces (TY+1) x TX
constexpr int SLOT_I = TY * (TX + 1);
constexpr int SLOT_J = (TY + 1) * TX;
constexpr int EX = SLOT_I + SLOT_J;

template <int NV, bool VISC>
struct Rec : RecF<NV, VISC> {   // staged-plane and private-slot sizes of this tile shape
  using RecF<NV, VISC>::NR;
  static constexpr int PLANE = NV * PSQ + NR * PS;         // doubles per staged plane
  // thread-private slots of the main threads ([field][NMAIN]): carried k-direction state hi_k (NV) and F_k (NV+3), norm
  // partials (NV+1), q of plane k+2 (NV), volume of planes k / k+1 (2).  Kept out of registers so that the i/j phases have room
  // for interleaved dependency chains (a DFMA has 8.4 cycles of latency, the pipe takes one per 2.1).
  static constexpr int OFF_PHI = 0, OFF_PF = NV, OFF_PN = 2 * NV + 3, OFF_PQ2 = 3 * NV + 4, OFF_PVOL = 4 * NV + 4, NPRIV = 4 * NV + 6;
  static constexpr int SMEM = 2 * PLANE + NV * EX + (NV + 3) * EX + NPRIV * NMAIN;
};


template <int NV, int INTERP, int SCHEME, bool VISC>
__global__ void __launch_bounds__(NT, 1) k_sweep(const Params P, const KArgs a) {
  using RC = Rec<NV, VISC>;
  constexpr bool SST = (NV == 7);
  constexpr bool SMQ = (INTERP == F3D_MUSCL || INTERP == F3D_INTERP_NONE);   // 3-point stencils read the staged planes
  extern __shared__ double smem[];
  double* const sm_hi = smem + 2 * RC::PLANE;          // [NV][EX]
  double* const sm_F = sm_hi + NV * EX;                // [NV+3][EX]
  double* const priv = sm_F + (NV + 3) * EX + (threadIdx.x < NMAIN ? threadIdx.x : 0);   // [NPRIV][NMAIN], this thread's column
  const Layout& Ly = P.L;
  const long long fs = Ly.fs;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int i0 = 1 + blockIdx.x * TX, j0 = 1 + blockIdx.y * TY;
  const int kb = 1 + blockIdx.z * a.kchunk, ke = min(kb + a.kchunk, Ly.kmx);   // planes kb .. ke-1
  const bool need_dt = a.first_stage != 0;
  const bool flux_on_k = Ly.kmx != 2;   // H = 0 when kmx == 2 (ausm.f90:205-210)
  const bool k_active = flux_on_k || VISC || need_dt;

  // ---- role of this thread ------------------------------------------------------------------------------------------
  int i, j, s0;                     // cell of this thread in the plane and its staged slot
  int dirh = -1;                    // halo threads: the one direction they serve
  bool act, stg, own = false;       // act: reconstructs / evaluates faces; stg: stages its cell record (a superset)
  bool rec0 = false, rec1 = false, fac0 = false, fac1 = false, wr_hi = true;
  int sminus = 0, splus = 0;        // halo threads: staged slots of the two neighbours along dirh (one of them an outer slot)
  int xhi = 0, xlo = 0;             // exchange slots: where the hi value / flux is written, where L / the high-face flux is read
  long long outer_off = 0;          // global offset (relative to the own cell) of the outer neighbour a halo thread stages
  int outer_slot = 0;
  if (wid < TY) {
    const int tx = lane, ty = wid;
    i = i0 + tx; j = j0 + ty;
    own = (i <= Ly.imx - 1) && (j <= Ly.jmx - 1);
    rec0 = fac0 = (j <= Ly.jmx - 1) && (i <= Ly.imx);
    rec1 = fac1 = (i <= Ly.imx - 1) && (j <= Ly.jmx);
    act = rec0 || rec1;
    stg = (i <= Ly.imx + 1) && (j <= Ly.jmx + 1);   // the ghost cells next to the last faces feed the reconstruction there
    s0 = (ty + 1) * PW + tx + 1;
  } else if (wid == TY) {           // the two i columns next to the tile: lanes 0..TY-1 low side, TY..2TY-1 high side
    const int r = lane % TY, side = lane / TY;
    i = (side == 0) ? i0 - 1 : i0 + TX; j = j0 + r;
    act = (side < 2) && (j <= Ly.jmx - 1) && (i <= Ly.imx);
    stg = (side < 2) && (j <= Ly.jmx + 1) && (i <= Ly.imx + 1);
    dirh = 0; rec0 = act; fac0 = act && side == 1; wr_hi = side == 0;
    s0 = (r + 1) * PW + (side == 0 ? 0 : TX + 1);
    outer_slot = PS + side * TY + r;
    outer_off = (side == 0) ? -1 : 1;
    sminus = (side == 0) ? outer_slot : s0 - 1; splus = (side == 0) ? s0 + 1 : outer_slot;
    xhi = r * (TX + 1) + (side == 0 ? 0 : TX); xlo = r * (TX + 1) + TX;
  } else {                          // low (wid == TY+1) and high (wid == TY+2) j rows next to the tile
    const bool high = wid == TY + 2;
    i = i0 + lane; j = high ? j0 + TY : j0 - 1;
    act = (i <= Ly.imx - 1) && (j <= Ly.jmx);
    stg = (i <= Ly.imx + 1) && (j <= Ly.jmx + 1);
    dirh = 1; rec1 = act; fac1 = act && high; wr_hi = !high;
    s0 = (high ? TY + 1 : 0) * PW + lane + 1;
    outer_slot = PS + 2 * TY + (high ? TX : 0) + lane;
    outer_off = high ? Ly.sj : -Ly.sj;
    sminus = high ? s0 - PW : outer_slot; splus = high ? outer_slot : s0 + PW;
    xhi = SLOT_I + (high ? TY * TX : 0) + lane; xlo = SLOT_I + TY * TX + lane;
  }
  if (i > Ly.imx + 1) i = Ly.imx + 1;
  if (j > Ly.jmx + 1) j = Ly.jmx + 1;
  const bool main_thr = wid < TY;

  const double* __restrict__ q = a.q;
  const double* __restrict__ vol = a.geom + (long long)G_VOL * fs;

  // stage the record of this thread's cell at plane kk into ring buffer kk & 1 (asynchronously)
  auto stage = [&](int kk) {
    if (!stg) return;
    double* pl = smem + (kk & 1) * RC::PLANE;
    const long long c1 = Ly.idx(i, j, kk);
#pragma unroll
    for (int v = 0; v < NV; ++v) cp_async8(pl + v * PSQ + s0, q + v * fs + c1);
    if (!main_thr && SMQ && act) {
#pragma unroll
      for (int v = 0; v < NV; ++v) cp_async8(pl + v * PSQ + outer_slot, q + v * fs + c1 + outer_off);
    }
    if (VISC) {
      double* pr = pl + NV * PSQ;
#pragma unroll
      for (int f = 0; f < RC::NGF; ++f) cp_async8(pr + f * PS + s0, a.grad + f * fs + c1);
#pragma unroll
      for (int f = 0; f < RC::NMU; ++f) cp_async8(pr + (RC::OFF_MU + f) * PS + s0, a.mu + f * fs + c1);
#pragma unroll
      for (int f = 0; f < 3; ++f) cp_async8(pr + (RC::OFF_C + f) * PS + s0, a.geom + (long long)(G_CX + f) * fs + c1);
    }
    if (own) cp_async8(priv + (RC::OFF_PVOL + (kk & 1)) * NMAIN, vol + c1);
  };

  // carried along k (thread-private smem): hi_k = value at the high k face of the current plane, F_k = flux of its low k face
  if (main_thr) {
#pragma unroll
    for (int f = 0; f < RC::OFF_PQ2; ++f) priv[f * NMAIN] = 0.0;
  }

  if (main_thr) stage(kb - 1);
  PT_DECL

  for (int k = kb - 1; k < ke; ++k) {
    const bool inplane = k >= kb;
    const long long c = Ly.idx(i, j, k);
    double* const plA = smem + (k & 1) * RC::PLANE;          // plane k
    double* const plB = smem + ((k + 1) & 1) * RC::PLANE;    // plane k+1
    const double* const qA = plA + s0;                       // staged q of this thread's cell, field stride PSQ
    const double* const rA = plA + NV * PSQ + s0;            // its record, field stride PS
    PT_MARK(7)
    stage(k + 1);                                            // overlaps with the in-plane work below
    PT_MARK(0)
    if (own && k_active && SMQ) {
#pragma unroll
      for (int v = 0; v < NV; ++v) cp_async8(priv + (RC::OFF_PQ2 + v) * NMAIN, q + v * fs + c + 2 * Ly.sk);
    }

    // ---- i and j: reconstruct, exchange, flux -------------------------------------------------------------------------
#pragma unroll 1
    for (int d = 0; d < 2; ++d) {
      const bool dorec = inplane && (d == 0 ? rec0 : rec1), doface = inplane && (d == 0 ? fac0 : fac1);
      const int pos = (d == 0) ? i : j, mx = (d == 0) ? Ly.imx : Ly.jmx;
      const int nb = (d == 0) ? 1 : PW;                       // staged-slot stride of the direction
      int exw, exr;                                           // exchange slots (hi value written / L read; flux written)
      if (main_thr) {
        const int tx = lane, ty = wid;
        exw = (d == 0) ? ty * (TX + 1) + tx + 1 : SLOT_I + (ty + 1) * TX + tx;
        exr = (d == 0) ? ty * (TX + 1) + tx : SLOT_I + ty * TX + tx;
      } else { exw = xhi; exr = xlo; }
      double gA_ = 0.0, gnx = 0.0, gny = 0.0, gnz = 0.0;   // face metrics, requested before the reconstruction so their latency overlaps it
      if (doface) {
        const double* __restrict__ gp = a.geom + (long long)(G_IA + 4 * d) * fs + c;
        gA_ = gp[0]; gnx = gp[fs]; gny = gp[2 * fs]; gnz = gp[3 * fs];
      }
      double lo[NV];
      if (dorec) {
        double hi[NV];
        if (SMQ) {
          const int om = main_thr ? -nb : sminus - s0, op = main_thr ? nb : splus - s0;
          double qm[NV], q0[NV], qp[NV];
#pragma unroll
          for (int v = 0; v < NV; ++v) { qm[v] = qA[v * PSQ + om]; q0[v] = qA[v * PSQ]; qp[v] = qA[v * PSQ + op]; }
          recon3<NV, INTERP>(P, qm, q0, qp, pos, mx, d, hi, lo);
        } else {
          line_cell_values<NV, INTERP>(P, q, vol, c, (d == 0) ? 1 : Ly.sj, pos, mx, d, hi, lo);
        }
        if (wr_hi) {
#pragma unroll
          for (int v = 0; v < NV; ++v) sm_hi[v * EX + exw] = hi[v];
        }
      }
      PT_MARK(1)
      __syncthreads();
      PT_MARK(2)
      if (doface) {
        double L[NV], F[NV], lam = 0.0, vis = 0.0, tur = 0.0;
        const int exl = main_thr ? exr : ((d == 0) ? (exr - 0) : exr);   // L sits at the slot of the low neighbour's hi value
#pragma unroll
        for (int v = 0; v < NV; ++v) L[v] = sm_hi[v * EX + exl];
        face_eval<NV, SCHEME, VISC, PS, PSQ>(P, d, qA - nb, qA, rA - nb, rA, gA_, gnx, gny, gnz, pos, mx, L, lo, true, need_dt, F, lam, vis, tur);
#pragma unroll
        for (int v = 0; v < NV; ++v) sm_F[v * EX + exr] = F[v];
        if (need_dt) {
          sm_F[NV * EX + exr] = lam;
          if (VISC) sm_F[(NV + 1) * EX + exr] = vis;
          if (VISC && SST) sm_F[(NV + 2) * EX + exr] = tur;
        }
      }
      PT_MARK(3)
    }

    // ---- k: the plane k+1 record of this column has landed ---------------------------------------------------------------
    double kA = 0.0, knx = 0.0, kny = 0.0, knz = 0.0;
    if (own && k_active) {   // metrics of the k face, requested before the wait so their latency overlaps it
      const double* __restrict__ gp = a.geom + (long long)G_KA * fs + c + Ly.sk;
      kA = gp[0]; knx = gp[fs]; kny = gp[2 * fs]; knz = gp[3 * fs];
    }
    cp_async_wait_all();
    PT_MARK(4)
    double F_n[NV + 3];
#pragma unroll
    for (int v = 0; v < NV + 3; ++v) F_n[v] = 0.0;
    if (own && k_active) {
      const double* const qB = plB + s0;
      const double* const rB = plB + NV * PSQ + s0;
      double L[NV];
      if (k == kb - 1) {   // prime the carried hi value: cell kb-1 reconstructed along k
        double lo_[NV];
        if (SMQ) {
          double qm[NV], q0[NV], qp[NV];
#pragma unroll
          for (int v = 0; v < NV; ++v) { qm[v] = q[v * fs + c - Ly.sk]; q0[v] = qA[v * PSQ]; qp[v] = qB[v * PSQ]; }
          recon3<NV, INTERP>(P, qm, q0, qp, k, Ly.kmx, 2, L, lo_);
        } else {
          line_cell_values<NV, INTERP>(P, q, vol, c, Ly.sk, k, Ly.kmx, 2, L, lo_);
        }
      } else {
#pragma unroll
        for (int v = 0; v < NV; ++v) L[v] = priv[(RC::OFF_PHI + v) * NMAIN];
      }
      double lo[NV], hi_n[NV];
      if (SMQ) {
        double qm[NV], q0[NV], q2[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) { qm[v] = qA[v * PSQ]; q0[v] = qB[v * PSQ]; q2[v] = priv[(RC::OFF_PQ2 + v) * NMAIN]; }
        recon3<NV, INTERP>(P, qm, q0, q2, k + 1, Ly.kmx, 2, hi_n, lo);
      } else {
        line_cell_values<NV, INTERP>(P, q, vol, c + Ly.sk, Ly.sk, k + 1, Ly.kmx, 2, hi_n, lo);
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) priv[(RC::OFF_PHI + v) * NMAIN] = hi_n[v];
      double F[NV], lam = 0.0, vis = 0.0, tur = 0.0;
      face_eval<NV, SCHEME, VISC, PS, PSQ>(P, 2, qA, qB, rA, rB, kA, knx, kny, knz, k + 1, Ly.kmx, L, lo, flux_on_k, need_dt, F, lam, vis, tur);
#pragma unroll
      for (int v = 0; v < NV; ++v) F_n[v] = F[v];
      F_n[NV] = lam; F_n[NV + 1] = vis; F_n[NV + 2] = tur;
    }
    PT_MARK(5)
    __syncthreads();
    PT_MARK(2)

    // ---- the cell -----------------------------------------------------------------------------------------------------------
    if (inplane && own) {
      const int tx = lane, ty = wid;
      const int sl0 = ty * (TX + 1) + tx, sh0 = sl0 + 1, sl1 = SLOT_I + ty * TX + tx, sh1 = sl1 + TX;
      double res[NV];
      double merr = 0.0;
      double F_k[NV + 3];
#pragma unroll
      for (int v = 0; v < NV + 3; ++v) F_k[v] = priv[(RC::OFF_PF + v) * NMAIN];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const double Fl0 = sm_F[v * EX + sl0], Fh0 = sm_F[v * EX + sh0], Fl1 = sm_F[v * EX + sl1], Fh1 = sm_F[v * EX + sh1];
        double r = 0.0;
        r = r + (Fh0 - Fl0);   // scheme.f90:133-135
        r = r + (Fh1 - Fl1);
        if (k_active) r = r + (F_n[v] - F_k[v]);
        res[v] = r;
        if (v == 0) {          // resnorm.f90:190-198
          if (i == 1) merr += Fl0;
          if (i == Ly.imx - 1) merr -= Fh0;
          if (j == 1) merr += Fl1;
          if (j == Ly.jmx - 1) merr -= Fh1;
          if (k_active) {
            if (k == 1) merr += F_k[0];
            if (k == Ly.kmx - 1) merr -= F_n[0];
          }
        }
      }
      {
        bool bad = false;
#pragma unroll
        for (int v = 0; v < NV; ++v) bad |= isnan(res[v]);
        if (bad) flag_error(a.err, F3D_ERR_NAN_FLUX, i, j, k);
      }
      const double volc = priv[(RC::OFF_PVOL + (k & 1)) * NMAIN];
      double qc[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) qc[v] = qA[v * PSQ];
      if (SST && VISC) {   // source.f90:214-268
        double g[6][3];
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) {
          if (cc == 3) continue;
          g[cc][0] = rA[(3 * cc + 0) * PS]; g[cc][1] = rA[(3 * cc + 1) * PS]; g[cc][2] = rA[(3 * cc + 2) * PS];
        }
        const double mut = rA[(RC::OFF_MU + 1) * PS];
        const double density = qc[0], tk = qc[5], tw = qc[6];
        const double wx = g[2][1] - g[1][2], wy = g[0][2] - g[2][0], wz = g[1][0] - g[0][1];
        const double vort = sqrt(wx * wx + wy * wy + wz * wz);
        double CD = 2 * density * kSigmaW2 * (g[4][0] * g[5][0] + g[4][1] * g[5][1] + g[4][2] * g[5][2]) * rcp64(tw);
        CD = dmax(CD, P.cd_floor);
        const double F1 = rA[(RC::OFF_MU + 2) * PS];
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
          const double* lamv = sm_F + NV * EX;
          const double lmxsum = lamv[sl0] + lamv[sl1] + F_k[NV] + lamv[sh0] + lamv[sh1] + F_n[NV];
          dtc = rcp64(lmxsum);
          dtc = dtc * volc * P.CFL;
          if (VISC) {
            const double* visv = sm_F + (NV + 1) * EX;
            double s = visv[sl0] + visv[sl1] + F_k[NV + 1] + visv[sh0] + visv[sh1] + F_n[NV + 1];
            s = P.gm * s * P.inv_Pr;
            s = 2. * rcp64(s + (2. * P.CFL * volc * rcp64(dtc)));
            dtc = P.CFL * (s * volc);
            if (SST) {
              const double* turv = sm_F + (NV + 2) * EX;
              double t = turv[sl0] + turv[sl1] + F_k[NV + 2] + turv[sh0] + turv[sh1] + F_n[NV + 2];
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
        if (a.have_store || a.quse != a.q) {
          u1[0] = a.quse[c];
#pragma unroll
          for (int v = 1; v < NV; ++v) u1[v] = a.quse[v * fs + c] * u1[0];
        } else {   // the state the update starts from is the staged one
          u1[0] = qc[0];
#pragma unroll
          for (int v = 1; v < NV; ++v) u1[v] = qc[v] * u1[0];
        }
        u1[4] = (u1[4] * P.inv_gm1 + 0.5 * (u1[1] * u1[1] + u1[2] * u1[2] + u1[3] * u1[3])) * rcp64(u1[0]) + 0.;
        if (SST) {
          const double F1 = rA[(RC::OFF_MU + 2) * PS];
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
        priv[RC::OFF_PN * NMAIN] += merr;
#pragma unroll
        for (int v = 0; v < NV; ++v) priv[(RC::OFF_PN + 1 + v) * NMAIN] += res[v] * res[v];
      }
    }
    if (own && k_active) {
#pragma unroll
      for (int v = 0; v < NV + 3; ++v) priv[(RC::OFF_PF + v) * NMAIN] = F_n[v];
    }
    PT_MARK(6)
  }

  PT_FLUSH
  if (a.want_norms) {   // per-CTA partial: warp shuffle, then the block
    __syncthreads();
    double* sred = smem;   // [NV+1][NT/32]
#pragma unroll
    for (int v = 0; v <= NV; ++v) {
      double x = main_thr ? priv[(RC::OFF_PN + v) * NMAIN] : 0.0;
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

int sweep3_grid_ctas(const Layout& L);
int launch_sweep3(Ctx* ctx, KArgs& a);

// which generation of the sweep kernel runs: 3 (sweep3.cu) unless F3D_SWEEP_GEN=2 asks for the previous one (kept for A/B
// measurements of the same build: profiles/r01_g3_summary.md)
static int sweep_generation() {
  static int gen = 0;
  if (gen == 0) {
    const char* e = getenv("F3D_SWEEP_GEN");
    gen = (e && e[0] == '2') ? 2 : 3;
  }
  return gen;
}

int residual_grid_ctas(const Layout& L) {   // per-CTA norm partials: room for either generation
  const int chunk = pick_kchunk(L);
  const int g2 = ((L.imx - 1 + TX - 1) / TX) * ((L.jmx - 1 + TY - 1) / TY) * ((L.kmx - 1 + chunk - 1) / chunk);
  const int g3 = sweep3_grid_ctas(L);
  return g2 > g3 ? g2 : g3;
}

template <int NV, int INTERP, int SCHEME, bool VISC>
static int launch_one(Ctx* ctx, KArgs& a) {
  const Layout& L = ctx->P.L;
  a.kchunk = pick_kchunk(L);
  dim3 grid((L.imx - 1 + TX - 1) / TX, (L.jmx - 1 + TY - 1) / TY, (L.kmx - 1 + a.kchunk - 1) / a.kchunk);
  const size_t shm = sizeof(double) * Rec<NV, VISC>::SMEM;
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
  if (sweep_generation() == 3) rc = launch_sweep3(ctx, a);
  else if (ctx->P.sa || ctx->P.pb_switch[0] || ctx->P.pb_switch[1] || ctx->P.pb_switch[2]) rc = F3D_ERR_UNSUPPORTED;   // generation 3 only
  else if (ctx->P.viscous) rc = ctx->P.sst ? launch_interp<7, true>(ctx, a) : launch_interp<5, true>(ctx, a);
  else rc = launch_interp<5, false>(ctx, a);
  if (ctx->timing) cudaEventRecord(e1, ctx->stream);
  if (rc) return rc;
  F3D_CUDA(cudaGetLastError());
  return 0;
}

#ifdef F3D_PHASE_TIMING
extern "C" void fest3d_gpu_phase_dump() {
  unsigned long long h[16];
  cudaMemcpyFromSymbol(h, g_phase, sizeof(h));
  const char* nm[8] = {"stage issue", "recon i/j", "barrier wait", "face i/j", "cp.async wait", "k recon+face", "cell (update)", "loop top"};
  unsigned long long tot = 0;
  for (int n = 0; n < 8; ++n) tot += h[n];
  for (int n = 0; n < 8; ++n) printf("phase %-14s %6.2f %%\n", nm[n], 100.0 * h[n] / (double)tot);
  memset(h, 0, sizeof(h));
  cudaMemcpyToSymbol(g_phase, h, sizeof(h));
}
#endif

int launch_norms(Ctx* ctx, int slot) {
  const int nvp1 = ctx->P.L.nv + 1;
  k_norm_final<<<nvp1, 256, 0, ctx->stream>>>(ctx->red, ctx->red_blocks, nvp1, ctx->norms_dev + 1024, ctx->norms_dev + (long long)slot * nvp1);
  ctx->launches++;
  F3D_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace f3d
