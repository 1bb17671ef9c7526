// Residual / time-step / update kernel with STAGED gradients (generation 3, "direction-specialised warps"): the default viscous
// path (F3D_GRADIENTS=staged); the gradients and viscosities it stages come from grad.cu.  Inviscid runs and F3D_GRADIENTS=fused use
// fused_kernel.cuh.
//
// Same data flow as generation 2 (sweep.cu): one CTA owns a TX x TY column of cells and marches through a chunk of k
// planes, the per-cell records (q, Green-Gauss gradients, mu / mu_t / F1, centre) of planes k and k+1 are staged in shared
// memory by cp.async, every reconstruction and every face flux is evaluated once.  What changed is WHO does the work.
// Generation 2 ran all three directions in every "main" thread (188 registers, 8 warps per SM, 5 main warps on 4 SM
// sub-partitions, three CTA-wide barriers per plane: 32 % FP64 pipe, profiles/r01_g2_summary.md).  Here the three
// directions of the SAME tile are given to three different warp groups that run concurrently:
//
//   I rows   (TY warps)  reconstruct along i, exchange through smem (named barrier 1, I group only), i-face fluxes
//   J rows   (TY warps)  the same along j (named barrier 2)
//   K rows   (TY warps)  own the cells: k reconstruction and k-face flux with the k state carried privately, and the
//                        cell work (residual assembly, SST source, local time step, update, norm partials) of the
//                        PREVIOUS plane, whose i/j fluxes the other groups left in a double-buffered exchange area
//   3 halo warps         the i columns / j rows just outside the tile (as in generation 2)
//
// so an SM sub-partition holds one I, one J, one K warp (+ at most one halo warp): 3.75 warps per sub-partition instead
// of 2, each with a third of the live state (<= 136 registers), equal work per sub-partition, and ONE CTA-wide barrier
// per plane.  The groups only meet at that barrier; the i/j fluxes of plane k are consumed one iteration later.
//
// Reference pipeline reproduced: as listed in sweep.cu (src/update.f90:534-545, 228-491; src/face/state/*.f90;
// src/boundary/boundary_state_reconstruction.f90:93-131; src/face/flux/convective/*.f90, scheme.f90:111-141;
// src/viscous.f90:144-447; src/source.f90:158-270; src/time.f90:122-246,366-531; src/resnorm.f90:171-199).
#pragma once
// Staging: the planes are filled by the TMA engine -- three 4-D tensor copies per plane (cp.async.bulk.tensor -> UTMALDG: q,
// gradients, aux = mu fields + cell centre), completion on an mbarrier.  Measured against per-thread cp.async in the same call:
// 4.66 vs 5.51 ms per launch at 256^3 (sweep3_kernel_cpasync.cuh keeps that form; one cp.async.bulk per field row tied with it:
// profiles/r01_g3_summary.md).
#include "sweep_common.cuh"

namespace f3d {
namespace g3 {

#ifdef F3D_PHASE_TIMING   // development aid: clock64 totals per warp: work of a plane, wait at the plane barrier
__device__ unsigned long long g_phase3[16 * 2];
#define PT3_DECL unsigned long long pt_t = clock64(), pt_acc[2] = {0, 0};
#define PT3_MARK(n) { const unsigned long long t_ = clock64(); pt_acc[n] += t_ - pt_t; pt_t = t_; }
#define PT3_FLUSH if (lane == 0) { atomicAdd(&g_phase3[wid * 2], pt_acc[0]); atomicAdd(&g_phase3[wid * 2 + 1], pt_acc[1]); }
#else
#define PT3_DECL
#define PT3_MARK(n)
#define PT3_FLUSH
#endif

constexpr int TX = 32, TY = 4;
constexpr int NMAIN = TX * TY;
constexpr int NW = 3 * TY + 4;
constexpr int NT = 32 * NW;
constexpr int W_IH = 3 * TY, W_JH = 3 * TY + 1, W_JL = 3 * TY + 2, W_C = 3 * TY + 3;
constexpr int ROWS_JL = TY / 2;   // rows whose cell work the low-j halo warp does after its (short) reconstruction; W_C does the rest
constexpr int N_IGRP = 32 * (TY + 1), N_JGRP = 32 * (TY + 2);   // threads on named barriers 1 and 2

// Staged plane.  Rows of PW = TX+4 cells (i0-2 .. i0+TX+1: the tile, its ring and the second ring cell the halo threads'
// reconstruction reads), because that is what the bulk-copy engine can fetch: a row of a field is contiguous in HBM, starts
// at an even element index (16-byte aligned) at i0-2, and 36 doubles are a multiple of 16 bytes.  Record fields (gradients,
// mu / mu_t / F1, centre) hold TY+2 rows (j0-1 .. j0+TY), q fields TY+4 rows (j0-2 .. j0+TY+1).  A "slot" is the record index
// s = row*PW + col with row 0 = j0-1; the same cell of a q field sits at s + PW.
constexpr int PW = TX + 4;
constexpr int PS = PW * (TY + 2);
constexpr int PSQ = PW * (TY + 4);
// exchange area ([field][slot], slot = face): i faces TY x (TX+1), j faces (TY+1) x TX
constexpr int SLOT_I = TY * (TX + 1);
constexpr int SLOT_J = (TY + 1) * TX;
constexpr int EX = SLOT_I + SLOT_J;

template <int NV, bool VISC>
struct Sm : RecF<NV, VISC> {
  using RecF<NV, VISC>::NR;
  static constexpr int NF = NV + 3;                       // flux + the lambda / viscous / turbulent face terms of the time step
  static constexpr int PLANE = NV * PSQ + NR * PS;        // doubles per staged plane
  static constexpr int OFF_X = 2 * PLANE;                 // exchange area [2][NF][EX]: hi values, then fluxes (same slot)
  static constexpr int NPK = (NV == 6 || NV == 8) ? 5 : 4;   // cell packet: volume; sst: F1, S_k, S_w [, S_gamma]; sa: vorticity, S_v, mu, dist
  static constexpr int OFF_PK = OFF_X + 2 * NF * EX;      // cell packets [2][NPK][NMAIN], written by the I rows
  static constexpr int OFF_PRIV = OFF_PK + 2 * NPK * NMAIN;   // private slots of the K threads, [field][NMAIN]:
  static constexpr int P_FK = 0;                          //   [3][NF] k-face flux; the face below plane p sits in third p % 3
  static constexpr int P_HI = 3 * NF;                     //   [2][NV] value at the high k face of the cell of plane p: half p & 1
  static constexpr int P_Q2 = P_HI + 2 * NV;              //   [NV] q of plane k+2 (not with eight variables: no room, read from global memory)
  static constexpr int P_VOL = P_Q2 + (NV == 8 ? 0 : NV); //   [2] volume of planes (p & 1)
  static constexpr int NPRIV = P_VOL + 2;
  static constexpr int OFF_NRM = OFF_PRIV + NPRIV * NMAIN;   // norm partials of the 64 threads that do cell work, [NV+1][64]
  static constexpr int OFF_MBAR = OFF_NRM + (NV + 1) * 64;   // two mbarriers (one per staged-plane buffer)
  static constexpr int TOTAL = OFF_MBAR + 2;
};

// cp.async.bulk (TMA engine, UBLKCP) + mbarrier: one row of one field per copy, completion counted in bytes on the mbarrier
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* b, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, void* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(b))
               : "memory");
}
__device__ __forceinline__ void tma_g2s_4d(void* dst, const CUtensorMap* tm, int x, int y, int z, int f, void* b) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(dst)),
               "l"(tm), "r"(x), "r"(y), "r"(z), "r"(f), "r"(smem_u32(b))
               : "memory");
}
struct TMaps { CUtensorMap q, grad, aux; };
__device__ __forceinline__ void bar_all() { asm volatile("bar.sync 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const double* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void bar_group(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// Cell work of one cell (row r of the tile, plane kc): residual assembly from the six face fluxes, SST source, local time
// step, point-implicit k/omega scaling, RK accumulation, conservative update, norm partials.  Everything it needs was left in
// shared memory during the previous iteration: i/j fluxes (exchange half kc & 1), k fluxes (thirds kc % 3 and (kc+1) % 3 of
// the K threads' ring), and the cell packet of the I rows (q, volume, F1, source terms).
template <int NV, bool VISC>
__device__ __forceinline__ void cell_work(const Params& P, const KArgs& a, double* __restrict__ smem, int tx, int r, int i, int j, int kc,
                                          bool need_dt, bool k_active, double* __restrict__ nrm /* [NV+1], stride 64 */, bool kkl = false) {
  using S = Sm<NV, VISC>;
  constexpr bool SST = (NV >= 7), LCTM = (NV == 8), SA = (NV == 6), TURB = SST || SA;
  constexpr int NF = S::NF;
  const Layout& Ly = P.L;
  const long long fs = Ly.fs;
  const long long cc = Ly.idx(i, j, kc);
  const int cell = r * TX + tx;
  const int sl0 = r * (TX + 1) + tx, sh0 = sl0 + 1, sl1 = SLOT_I + r * TX + tx, sh1 = sl1 + TX;
  const double* const xF = smem + S::OFF_X + (kc & 1) * NF * EX;                                  // i/j face fluxes of plane kc
  const double* const pk = smem + S::OFF_PK + (kc & 1) * S::NPK * NMAIN + cell;                   // cell packet
  const double* const Flo = smem + S::OFF_PRIV + (S::P_FK + (kc % 3) * NF) * NMAIN + cell;        // k face below the cell
  const double* const Fhi = smem + S::OFF_PRIV + (S::P_FK + ((kc + 1) % 3) * NF) * NMAIN + cell;  // k face above it
  double qc[NV];   // the state of the cell: read again from global memory (L2: it was staged two planes ago)
#pragma unroll
  for (int v = 0; v < NV; ++v) qc[v] = a.q[v * fs + cc];
  const double volc = pk[0];
  double res[NV];
  double merr = 0.0;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const double Fl0 = xF[v * EX + sl0], Fh0 = xF[v * EX + sh0], Fl1 = xF[v * EX + sl1], Fh1 = xF[v * EX + sh1];
    double rr = 0.0;
    rr = rr + (Fh0 - Fl0);   // scheme.f90:133-135
    rr = rr + (Fh1 - Fl1);
    if (k_active) rr = rr + (Fhi[v * NMAIN] - Flo[v * NMAIN]);
    res[v] = rr;
    if (v == 0) {          // resnorm.f90:190-198
      if (i == 1) merr += Fl0;
      if (i == Ly.imx - 1) merr -= Fh0;
      if (j == 1) merr += Fl1;
      if (j == Ly.jmx - 1) merr -= Fh1;
      if (k_active) {
        if (kc == 1) merr += Flo[0];
        if (kc == Ly.kmx - 1) merr -= Fhi[0];
      }
    }
  }
  {
    bool bad = false;
#pragma unroll
    for (int v = 0; v < NV; ++v) bad |= isnan(res[v]);
    if (bad) flag_error(a.err, F3D_ERR_NAN_FLUX, i, j, kc);
  }
  if (SST && VISC) {
    res[5] = res[5] - pk[2 * NMAIN];
    res[6] = res[6] - pk[3 * NMAIN];
    if (LCTM) res[7] = res[7] - pk[4 * NMAIN];
  }
  if (SA && VISC) res[5] = res[5] - pk[2 * NMAIN];

  double dtc = 0.0;
  if (need_dt) {
    if (P.time_stepping == 1 && P.global_time_step > 0) {
      dtc = P.global_time_step;
    } else {
      const double* lamv = xF + NV * EX;
      const double lmxsum = lamv[sl0] + lamv[sl1] + Flo[NV * NMAIN] + lamv[sh0] + lamv[sh1] + Fhi[NV * NMAIN];
      dtc = rcp64(lmxsum);
      dtc = dtc * volc * P.CFL;
      if (VISC) {
        const double* visv = xF + (NV + 1) * EX;
        double s = visv[sl0] + visv[sl1] + Flo[(NV + 1) * NMAIN] + visv[sh0] + visv[sh1] + Fhi[(NV + 1) * NMAIN];
        s = P.gm * s * P.inv_Pr;
        s = 2. * rcp64(s + (2. * P.CFL * volc * rcp64(dtc)));
        dtc = P.CFL * (s * volc);
        if (TURB) {
          const double* turv = xF + (NV + 2) * EX;
          double tt = turv[sl0] + turv[sl1] + Flo[(NV + 2) * NMAIN] + turv[sh0] + turv[sh1] + Fhi[(NV + 2) * NMAIN];
          tt = P.gm * tt * P.inv_tPr;
          tt = 2. * rcp64(tt + (2. * P.CFL * volc * rcp64(dtc)));
          dtc = P.CFL * (tt * volc);
        }
      }
    }
    a.dt[cc] = dtc;
  } else if (a.mode == MODE_UPDATE) {
    dtc = a.dt[cc];
  }

  if (a.mode == MODE_RESIDUE_ONLY) {
#pragma unroll
    for (int v = 0; v < NV; ++v) a.residue[v * fs + cc] = res[v];
  } else {   // update.f90:371-485
    double u1[NV], R[NV], u2[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) R[v] = res[v];
    if (a.have_store || a.quse != a.q) {
      u1[0] = a.quse[cc];
#pragma unroll
      for (int v = 1; v < NV; ++v) u1[v] = a.quse[v * fs + cc] * u1[0];
    } else {   // the state the update starts from is the staged one
      u1[0] = qc[0];
#pragma unroll
      for (int v = 1; v < NV; ++v) u1[v] = qc[v] * u1[0];
    }
    u1[4] = (u1[4] * P.inv_gm1 + 0.5 * (u1[1] * u1[1] + u1[2] * u1[2] + u1[3] * u1[3])) * rcp64(u1[0]) + 0.;
    if (SST && kkl) {   // update.f90:399-404: u1(6), u1(7) are rho*k, rho*kL here, used where the model has k, kL -- reproduced
      const double mu_c = pk[NMAIN], dist_c = a.geom[(long long)G_DIST * fs + cc];
      const double eta = u1[0] * dist_c * (sqrt(0.3 * u1[5]) / (20 * mu_c));
      const double fphi = (1 + kKklCd1 * eta) / (1 + (eta * eta) * (eta * eta));
      R[5] = R[5] / (1. + ((2.5 * (kKklCmu75 * sqrt(u1[0]) * (u1[5] * sqrt(u1[5])) / fmax(u1[6], 1.e-20)) + (2 * mu_c / (dist_c * dist_c))) * dtc));
      R[6] = R[6] / (1. + (6 * mu_c * fphi / (dist_c * dist_c)) * dtc);
    } else if (SST) {
      const double F1 = VISC ? pk[NMAIN] : 0.0;
      const double beta = kBeta1 * F1 + (1. - F1) * kBeta2;
      R[5] = R[5] * rcp64(1 + (beta * qc[6] * dtc));
      R[6] = R[6] * rcp64(1 + (2 * beta * qc[6] * dtc));
    }
    if (SA && VISC) {   // update.f90:405-420: u1(6) is rho*tv here, used where the model has tv -- reproduced
      const double vort = pk[NMAIN], mu_c = pk[3 * NMAIN], dist_c = pk[4 * NMAIN];
      const double kd = kKappaSA * dist_c, kd2 = kd * kd;
      const double xi = u1[5] * qc[0] / mu_c;
      const double fv1 = pow3(xi) / (pow3(xi) + pow3(kCv1));
      const double fv2 = 1.0 - xi / (1 + xi * fv1);
      const double scap = vort + u1[5] * fv2 / (kd2);
      const double rsa = fmin(u1[5] / (scap * kd2), 10.0);
      const double fw = sa_fw(rsa);
      R[5] = R[5] / (1. + ((-1.0 * u1[0] * kCb1 * scap) + (2.0 * u1[0] * kCw1 * fw * u1[5] / (dist_c * dist_c))) * dtc);
    }
    if (a.have_store) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const double rn = a.rstore[v * fs + cc] + a.SF * R[v];
        a.rstore[v * fs + cc] = rn;
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
      flag_error(a.err, F3D_ERR_NEGATIVE_STATE, i, j, kc);
#pragma unroll
      for (int v = 0; v < NV; ++v) a.qnew[v * fs + cc] = qc[v];
    } else {
#pragma unroll
      for (int v = 0; v < 5; ++v) a.qnew[v * fs + cc] = u2[v];
      if (SST) {
        a.qnew[5 * fs + cc] = (u2[5] >= 0.) ? u2[5] : qc[5];
        a.qnew[6 * fs + cc] = (u2[6] >= 0.) ? u2[6] : qc[6];
        // lctm2015: update_with (update.f90:349-362, 462-484) writes qp(1:5), qp(6) and qp(7) back and nothing else -- the explicit
        // integrators never advance the intermittency (only plusgs.f90:2129 does); its residual still enters R_store and the norms
        if (LCTM) a.qnew[7 * fs + cc] = qc[7];
      }
      if (SA) a.qnew[5 * fs + cc] = fmax(u2[5], 1.e-12);   // update.f90:474-475
    }
  }
  if (a.want_norms) {   // resnorm.f90:187-198
    nrm[0] += merr;
#pragma unroll
    for (int v = 0; v < NV; ++v) nrm[(1 + v) * 64] += res[v] * res[v];
  }
}

// RARE gates the code of the seldom-used options (pressure-based switching, transition = bc): compiled into a second set of
// instantiations (sweep3_rare.cu) because even switched off it cost the register-tight common path 4.5 % (5.82 vs 5.57 ms).
static_assert(TX == kG3TX && TY == kG3TY, "tensor-map boxes are encoded for this tile (api.cu)");
template <int NV, int INTERP, int SCHEME, bool VISC, bool RARE>
__global__ void __launch_bounds__(NT, 1) k_sweep3(const Params P, const KArgs a, const __grid_constant__ TMaps tm) {
  using S = Sm<NV, VISC>;
  constexpr bool SST = (NV >= 7), LCTM = (NV == 8), SA = (NV == 6), TURB = SST || SA;
  constexpr bool SMQ = (INTERP == F3D_MUSCL || INTERP == F3D_INTERP_NONE);   // 3-point stencils read the staged planes
  constexpr int NF = S::NF;
  extern __shared__ __align__(128) double smem[];
  const Layout& Ly = P.L;
  const long long fs = Ly.fs;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int i0 = 1 + blockIdx.x * TX, j0 = 1 + blockIdx.y * TY;
  const int kb = 1 + blockIdx.z * a.kchunk, ke = min(kb + a.kchunk, Ly.kmx);   // planes kb .. ke-1
  const bool need_dt = a.first_stage != 0;
  const bool flux_on_k = Ly.kmx != 2;   // H = 0 when kmx == 2 (ausm.f90:205-210)
  const bool k_active = flux_on_k || VISC || need_dt;
  const double* __restrict__ q = a.q;
  const double* __restrict__ vol = a.geom + (long long)G_VOL * fs;
  if (tid < 64) {       // norm partials of the threads that do cell work
#pragma unroll
    for (int v = 0; v <= NV; ++v) smem[S::OFF_NRM + v * 64 + tid] = 0.0;
  }

  // ---- role of this thread.  The values are re-derived from the thread index at the top of every plane (the empty asm keeps
  // the compiler from hoisting them out of the loop and spilling them: that cost 27 % of all stall samples during bring-up).
  // All flux warps -- I rows, J rows, K rows, halo warps -- then run ONE instruction stream, parameterised by these values:
  // separate code copies per direction made the SM fetch-bound (24 % no_instruction stalls, profiles/r01_g3_summary.md).
  int i, j, s0, d, cell, r_lo, r_hi;
  bool rec, fac, wr_hi, irow, krow, own;
  int om, op;                       // I/J: staged-slot offsets of the two neighbours along d
  int exw, exr;                     // I/J: exchange slots: where the hi value goes; where L is read and the flux written
  int pos, mx;
  auto role = [&]() {
    int t_ = tid;
    asm volatile("" : "+r"(t_));
    const int ln = t_ & 31, w = t_ >> 5;
    wr_hi = true; irow = false; krow = false; own = false; cell = 0; r_lo = 0; r_hi = 0;
    rec = fac = false; d = 0; i = i0 + ln; j = j0; s0 = PW + 2; om = op = 0; exw = exr = 0;
    if (w < TY) {                   // I row
      const int tx = ln, ty = w;
      d = 0; i = i0 + tx; j = j0 + ty; irow = true; cell = ty * TX + tx;
      rec = fac = (j <= Ly.jmx - 1) && (i <= Ly.imx);
      s0 = (ty + 1) * PW + tx + 2; om = -1; op = 1;
      exw = ty * (TX + 1) + tx + 1; exr = ty * (TX + 1) + tx;
    } else if (w < 2 * TY) {        // J row
      const int tx = ln, ty = w - TY;
      d = 1; i = i0 + tx; j = j0 + ty;
      rec = fac = (i <= Ly.imx - 1) && (j <= Ly.jmx);
      s0 = (ty + 1) * PW + tx + 2; om = -PW; op = PW;
      exw = SLOT_I + (ty + 1) * TX + tx; exr = SLOT_I + ty * TX + tx;
    } else if (w < 3 * TY) {        // K row: the column of its cell
      const int tx = ln, ty = w - 2 * TY;
      d = 2; i = i0 + tx; j = j0 + ty; krow = true; cell = ty * TX + tx;
      own = (i <= Ly.imx - 1) && (j <= Ly.jmx - 1);
      rec = fac = own && k_active;
      s0 = (ty + 1) * PW + tx + 2;
    } else if (w == W_IH) {         // the two i columns next to the tile: lanes 0..TY-1 low side, TY..2TY-1 high side
      const int r = ln % TY, side = ln / TY;
      d = 0; i = (side == 0) ? i0 - 1 : i0 + TX; j = j0 + r;
      rec = (side < 2) && (j <= Ly.jmx - 1) && (i <= Ly.imx);
      fac = rec && side == 1; wr_hi = side == 0;
      s0 = (r + 1) * PW + (side == 0 ? 1 : TX + 2); om = -1; op = 1;
      exw = r * (TX + 1) + (side == 0 ? 0 : TX); exr = r * (TX + 1) + TX;
    } else if (w == W_C) {          // cell work only
      r_lo = ROWS_JL; r_hi = TY;
    } else {                        // high (W_JH) and low (W_JL) j rows next to the tile
      const bool high = w == W_JH;
      d = 1; i = i0 + ln; j = high ? j0 + TY : j0 - 1;
      rec = (i <= Ly.imx - 1) && (j <= Ly.jmx);
      fac = rec && high; wr_hi = !high;
      s0 = (high ? TY + 1 : 0) * PW + ln + 2; om = -PW; op = PW;
      exw = SLOT_I + (high ? TY * TX : 0) + ln; exr = SLOT_I + TY * TX + ln;
      if (!high) { r_lo = 0; r_hi = ROWS_JL; }
    }
    if (i > Ly.imx + 1) i = Ly.imx + 1;
    if (j > Ly.jmx + 1) j = Ly.jmx + 1;
    pos = (d == 0) ? i : j; mx = (d == 0) ? Ly.imx : ((d == 1) ? Ly.jmx : Ly.kmx);
  };

  // ---- staging of a plane by the TMA engine: one elected thread issues a tensor copy per staged array (box = 36 columns x
  // TY+4 / TY+2 rows x 1 plane x all fields, landing as [field][row][col]); cells outside the arrays are zero-filled, so the byte
  // count the mbarrier expects is always that of the full boxes.  Tensor coordinates: x = i + 15, y = j + 2, z = k + 2
  // (ctx.hpp:Layout: idx = 15 + i + sj (j+2) + sk (k+2)).
  void* const mbar = smem + S::OFF_MBAR;
  constexpr unsigned plane_bytes = 8u * PW * ((TY + 4) * NV + (TY + 2) * S::NR);
  auto stage_plane = [&](int kk) {
    if (tid != NT - 1) return;
    double* const pl = smem + (kk & 1) * S::PLANE;
    char* const mb = (char*)mbar + 8 * (kk & 1);
    mbar_expect_tx(mb, plane_bytes);
    tma_g2s_4d(pl, &tm.q, i0 + 13, j0, kk + 2, 0, mb);
    if (VISC) {
      tma_g2s_4d(pl + NV * PSQ, &tm.grad, i0 + 13, j0 + 1, kk + 2, 0, mb);
      tma_g2s_4d(pl + NV * PSQ + S::NGFS * PS, &tm.aux, i0 + 13, j0 + 1, kk + 2, 0, mb);
    }
  };
  // parity of the mbarrier phase that completes when plane kk has landed (buffer kk & 1 is used by every second plane)
  auto plane_parity = [&](int kk) { return (unsigned)(((kk - (kb - 1)) >> 1) & 1); };
  if (tid == 0) {
    mbar_init(mbar, 1); mbar_init((char*)mbar + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  stage_plane(kb - 1);

  role();
  if (krow) {   // K rows: clear the k-face ring, stage plane kb-1, prime the carried value of cell kb-1 (from global memory)
    double* const priv = smem + S::OFF_PRIV + cell;
#pragma unroll
    for (int f = 0; f < S::P_Q2; ++f) priv[f * NMAIN] = 0.0;
    if (own) cp_async8(priv + (S::P_VOL + ((kb - 1) & 1)) * NMAIN, vol + Ly.idx(i, j, kb - 1));
    if (rec) {
      const long long c = Ly.idx(i, j, kb - 1);
      double L[NV], lo_[NV];
      if (SMQ) {
        double qm[NV], q0[NV], qp[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) { qm[v] = q[v * fs + c - Ly.sk]; q0[v] = q[v * fs + c]; qp[v] = q[v * fs + c + Ly.sk]; }
        const double p_far = (RARE && INTERP == F3D_MUSCL && P.pb_switch[2] && kb - 1 == 0) ? q[4 * fs + c + 2 * Ly.sk] : 0.0;
        recon3<NV, INTERP, RARE>(P, qm, q0, qp, kb - 1, Ly.kmx, 2, L, lo_, p_far);
      } else {
        line_cell_values<NV, INTERP, RARE>(P, q, vol, c, Ly.sk, kb - 1, Ly.kmx, 2, L, lo_);
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) priv[(S::P_HI + ((kb - 1) & 1) * NV + v) * NMAIN] = L[v];
    }
    cp_async_wait_all();
  }

  PT3_DECL
  for (int k = kb - 1; k <= ke; ++k) {
    PT3_MARK(0)
    bar_all();   // plane k is staged; the fluxes and cell packets of plane k-1 are complete
    PT3_MARK(1)
    role();
    // ---- staging of plane k+1 over plane k-1 (nobody reads plane k-1 any more: the cell work takes what it needs from the cell
    // packets); the K rows also fetch the volume of plane k+1 and their own q of plane k+2 for the k stencil
    if (k + 1 <= ke) stage_plane(k + 1);
    if (krow && k <= ke - 1) {
      if (own) cp_async8(smem + S::OFF_PRIV + (S::P_VOL + ((k + 1) & 1)) * NMAIN + cell, vol + Ly.idx(i, j, k + 1));
      if (rec && SMQ && !LCTM) {
        const long long c2 = Ly.idx(i, j, k + 2);
#pragma unroll
        for (int v = 0; v < NV; ++v) cp_async8(smem + S::OFF_PRIV + (S::P_Q2 + v) * NMAIN + cell, q + v * fs + c2);
      }
    }

    // ---- flux work: one reconstruction and one face per thread, the same code for every direction --------------------------------
    const bool active = (wid != W_C) && (k <= ke - 1) && (krow || k >= kb);
    if (active) {
      // I/J: cell (i,j,k), face below it along d.  K: cell (i,j,k+1), face between planes k and k+1.
      const int pA = (k & 1) * S::PLANE, pB = ((k + 1) & 1) * S::PLANE;
      const int nb = (d == 0) ? 1 : PW;
      const int o_m = krow ? pA + PW + s0 : pA + PW + s0 + om;       // stencil: low neighbour, cell, high neighbour (q field 0)
      const int o_0 = krow ? pB + PW + s0 : pA + PW + s0;
      const int o_p = krow ? S::OFF_PRIV + S::P_Q2 * NMAIN + cell : pA + PW + s0 + op;
      const int f_p = krow ? NMAIN : PSQ;                            // field stride of the high neighbour
      const int o_ql = krow ? pA + PW + s0 : pA + PW + s0 - nb;      // the two cells of the face (q field 0); their records sit at
                                                                     // the same offset minus PW plus NV*PSQ
      const int o_qh = o_0;
      const int f_x = krow ? NMAIN : EX;                             // field stride of the hi / L / flux slots
      const int o_hw = krow ? S::OFF_PRIV + (S::P_HI + ((k + 1) & 1) * NV) * NMAIN + cell : S::OFF_X + (k & 1) * NF * EX + exw;
      const int o_lr = krow ? S::OFF_PRIV + (S::P_HI + (k & 1) * NV) * NMAIN + cell : S::OFF_X + (k & 1) * NF * EX + exr;
      const int o_fw = krow ? S::OFF_PRIV + (S::P_FK + ((k + 1) % 3) * NF) * NMAIN + cell : o_lr;
      const int cpos = krow ? k + 1 : pos;                           // index of the cell and of the face along d
      const long long c = Ly.idx(i, j, k);
      const long long cg = krow ? c + Ly.sk : c;                     // global index of the cell / face
      double gA_ = 0.0, gnx = 0.0, gny = 0.0, gnz = 0.0;   // face metrics, requested before the reconstruction
      if (fac) {
        const double* __restrict__ gp = a.geom + (long long)(G_IA + 4 * d) * fs + cg;
        gA_ = gp[0]; gnx = gp[fs]; gny = gp[2 * fs]; gnz = gp[3 * fs];
      }
      if (krow) {   // the K rows read plane k+1 (and, in the first iteration, plane kb-1): wait until the bulk copies have landed
        cp_async_wait_all();
        mbar_wait((char*)mbar + 8 * (k & 1), plane_parity(k));
        mbar_wait((char*)mbar + 8 * ((k + 1) & 1), plane_parity(k + 1));
      } else {
        mbar_wait((char*)mbar + 8 * (k & 1), plane_parity(k));
      }
      double lo[NV];
      if (rec) {
        double hi[NV];
        if (SMQ) {
          double qm[NV], q0[NV], qp[NV];
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            qm[v] = smem[o_m + v * PSQ]; q0[v] = smem[o_0 + v * PSQ];
            qp[v] = (LCTM && krow) ? q[v * fs + cg + Ly.sk] : smem[o_p + v * f_p];   // plane k+2: no private slot with eight variables (Sm)
          }
          double p_far = 0.0;   // pressure-based switching at the two ghost positions reads the pressure two cells inwards
          if (RARE && INTERP == F3D_MUSCL && P.pb_switch[d] && (cpos == 0 || cpos == mx)) {
            const int two = (cpos == 0) ? 2 : -2;
            p_far = krow ? q[4 * fs + cg + two * Ly.sk] : smem[o_0 + 4 * PSQ + two * nb];
          }
          recon3<NV, INTERP, RARE>(P, qm, q0, qp, cpos, mx, d, hi, lo, p_far);
        } else {   // five-point stencils from global memory / L2.  Measured (profiles/r02_summary.md): taking them from the staged plane
                   // for the I / J row warps, whose stencils lie inside it, changes nothing (9.81 vs 9.63 ms): the WENO weights bind
          line_cell_values<NV, INTERP, RARE>(P, q, vol, cg, (d == 0) ? 1 : ((d == 1) ? Ly.sj : Ly.sk), cpos, mx, d, hi, lo);
        }
        if (wr_hi) {
#pragma unroll
          for (int v = 0; v < NV; ++v) smem[o_hw + v * f_x] = hi[v];
        }
      }
      if (!krow) bar_group(1 + d, (d == 0) ? N_IGRP : N_JGRP);
      if (fac) {
        double L[NV], F[NV], lam = 0.0, vis = 0.0, tur = 0.0;
#pragma unroll
        for (int v = 0; v < NV; ++v) L[v] = smem[o_lr + v * f_x];
        face_eval<NV, SCHEME, VISC, PS, PSQ>(P, d, smem + o_ql, smem + o_qh, smem + o_ql - PW + NV * PSQ, smem + o_qh - PW + NV * PSQ, gA_, gnx, gny, gnz, cpos, mx, L, lo,
                                             krow ? flux_on_k : true, need_dt, F, lam, vis, tur, RARE && P.kkl, &a,
                                             cg - ((d == 0) ? 1 : ((d == 1) ? Ly.sj : Ly.sk)), cg);
#pragma unroll
        for (int v = 0; v < NV; ++v) smem[o_fw + v * f_x] = F[v];
        if (need_dt || krow) {
          smem[o_fw + NV * f_x] = lam;
          if (VISC) smem[o_fw + (NV + 1) * f_x] = vis;
          if (VISC && TURB) smem[o_fw + (NV + 2) * f_x] = tur;
        }
      }
      if (irow && rec && i <= Ly.imx - 1) {   // I rows: the cell packet of the own cell for next iteration's cell work
        const double* const rA = smem + o_0 - PW + NV * PSQ;
        const double* const qA = smem + o_0;
        double* const pk = smem + S::OFF_PK + (k & 1) * S::NPK * NMAIN + cell;
        const double volc = smem[S::OFF_PRIV + (S::P_VOL + (k & 1)) * NMAIN + cell];
        pk[0] = volc;
        if (SST && VISC && RARE && P.kkl) {   // k-kL source terms (source.f90:607-832)
          const long long cI = c;
          const double density = qA[0], tk = qA[5 * PSQ], tkl = qA[6 * PSQ];
          double gv[3][3];   // velocity gradients of the cell: gv[component][direction]
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) { gv[cc][0] = rA[(3 * cc + 0) * PS]; gv[cc][1] = rA[(3 * cc + 1) * PS]; gv[cc][2] = rA[(3 * cc + 2) * PS]; }
          const double mut = rA[(S::OFF_MU + 1) * PS], mu_c = rA[S::OFF_MU * PS];
          const double S11 = 0.5 * (gv[0][0] + gv[0][0]), S12 = 0.5 * (gv[0][1] + gv[1][0]), S13 = 0.5 * (gv[0][2] + gv[2][0]);
          const double S22 = 0.5 * (gv[1][1] + gv[1][1]), S23 = 0.5 * (gv[1][2] + gv[2][1]), S33 = 0.5 * (gv[2][2] + gv[2][2]);
          const double delv = gv[0][0] + gv[1][1] + gv[2][2];
          const double tkk = (2.0 / 3.0) * density * tk;
          const double T11 = mut * (2 * S11 - (2.0 / 3.0) * delv) - tkk, T22 = mut * (2 * S22 - (2.0 / 3.0) * delv) - tkk, T33 = mut * (2 * S33 - (2.0 / 3.0) * delv) - tkk;
          const double T12 = mut * (2 * S12), T13 = mut * (2 * S13), T23 = mut * (2 * S23);
          double P_k = 0.;
          P_k = P_k + T11 * gv[0][0] + T12 * gv[0][1] + T13 * gv[0][2];
          P_k = P_k + T12 * gv[1][0] + T22 * gv[1][1] + T23 * gv[1][2];
          P_k = P_k + T13 * gv[2][0] + T23 * gv[2][1] + T33 * gv[2][2];
          const double D_k = kKklCmu75 * density * ((tk * tk) * sqrt(tk)) / fmax(tkl, 1.e-20);
          P_k = fmin(P_k, 20 * D_k);
          // |d2u| of the von Karman length scale: made by k_kkl_udd (grad.cu) from the finished gradient arrays and staged in the aux slot
          // the SST models use for F1 (k-kL has no blending function).  In this packet it cost the I-row warps 27 neighbour reads of plane
          // k-1 / k+1 and 18 face metrics from global memory while the other warps waited.
          const double udd = rA[(S::OFF_MU + 2) * PS];
          const double ud = sqrt(2 * (S11 * S11 + S12 * S12 + S13 * S13 + S12 * S12 + S22 * S22 + S23 * S23 + S13 * S13 + S23 * S23 + S33 * S33));
          const double dist_c = a.geom[(long long)G_DIST * fs + cI];
          double Lvk = kKklKappa * fabs(ud / fmax(udd, 1.e-20));
          const double fp = fmin(fmax(P_k / D_k, 0.5), 1.0);
          Lvk = fmax(Lvk, tkl / fmax((tk * kKklC11), 1.e-20));
          Lvk = fmin(Lvk, kKklC12 * kKklKappa * dist_c * fp);
          const double eta = density * dist_c * sqrt(0.3 * tk) / (20 * mu_c);
          const double fphi = (1 + kKklCd1 * eta) / (1 + (eta * eta) * (eta * eta));
          const double rr = (tkl / fmax(tk * Lvk, 1.e-20));
          const double cphi1 = (kKklZeta1 - kKklZeta2 * (rr * rr));
          const double P_kl = cphi1 * tkl * P_k / fmax(tk, 1.e-20);
          const double D_kl = kKklZeta3 * density * (tk * sqrt(tk));
          const double S_k = P_k - D_k - 2 * mu_c * tk / (dist_c * dist_c);
          const double S_kl = P_kl - D_kl - 6 * mu_c * tkl * fphi / (dist_c * dist_c);
          pk[NMAIN] = mu_c;
          pk[2 * NMAIN] = S_k * volc;
          pk[3 * NMAIN] = S_kl * volc;
        } else if (LCTM && VISC) {
          // SST + gamma-transition source terms (source.f90:273-463): made by k_gradients<7> (grad.cu:lctm_sources), where the gradients, mu_t and
          // F1 of the cell are in registers, and read here from global memory -- the packet (three exponentials, a dozen divisions) sat on the
          // I-row warps alone while the other warps waited
          pk[NMAIN] = rA[(S::OFF_MU + 2) * PS];
          pk[2 * NMAIN] = a.src[c];
          pk[3 * NMAIN] = a.src[fs + c];
          pk[4 * NMAIN] = a.src[2 * fs + c];
        } else if (SST && VISC) {   // SST source terms (source.f90:214-268)
          double g[6][3];
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) {
            if (cc == 3) continue;
            g[cc][0] = rA[(3 * cc + 0) * PS]; g[cc][1] = rA[(3 * cc + 1) * PS]; g[cc][2] = rA[(3 * cc + 2) * PS];
          }
          const double mut = rA[(S::OFF_MU + 1) * PS];
          const double F1c = rA[(S::OFF_MU + 2) * PS];
          const double density = qA[0], tk = qA[5 * PSQ], tw = qA[6 * PSQ];
          const double wx = g[2][1] - g[1][2], wy = g[0][2] - g[2][0], wz = g[1][0] - g[0][1];
          const double vort = sqrt(wx * wx + wy * wy + wz * wz);
          double CD = 2 * density * kSigmaW2 * (g[4][0] * g[5][0] + g[4][1] * g[5][1] + g[4][2] * g[5][2]) * rcp64(tw);
          CD = dmax(CD, P.cd_floor);
          const double gama = P.gama1 * F1c + P.gama2 * (1. - F1c);
          const double beta = kBeta1 * F1c + kBeta2 * (1. - F1c);
          const double D_k = kBstar * density * tw * tk;
          const double D_w = beta * density * (tw * tw);
          const double divergence = g[0][0] + g[1][1] + g[2][2];
          double P_k = mut * (vort * vort) - ((2.0 / 3.0) * density * tk * divergence);
          P_k = dmin(P_k, P.pk_limiter * D_k);
          double P_w = (density * gama * rcp64(mut)) * P_k;
          double lamda = (1. - F1c) * CD;
          if (RARE && P.trans_bc) {   // add_sst_bc_source (source.f90:467-604): no CD floor, P_k = mu_t vort^2 capped at 20 D_k, gamma_BC on P_k
            const double CDb = 2 * density * kSigmaW2 * (g[4][0] * g[5][0] + g[4][1] * g[5][1] + g[4][2] * g[5][2]) / tw;
            const double gam0 = P.gama1_default * F1c + P.gama2_default * (1. - F1c);
            P_k = fmin(mut * (vort * vort), 20.0 * D_k);
            P_w = (density * gam0 / mut) * P_k;
            lamda = (1. - F1c) * CDb;
            const double u_ = qA[PSQ], v_ = qA[2 * PSQ], w_ = qA[3 * PSQ];
            const double vmag = sqrt(((u_ * u_) + (v_ * v_)) + (w_ * w_));
            const double dist_c = a.geom[(long long)G_DIST * fs + c];
            const double mu_c = rA[S::OFF_MU * PS];
            const double re_v = density * dist_c * dist_c * vort / mu_c;
            P_k = gamma_bc(P.re_theta_t, P.nu_cr, mut / density, vmag, dist_c, re_v) * P_k;
          }
          pk[NMAIN] = F1c;
          pk[2 * NMAIN] = (P_k - D_k) * volc;
          pk[3 * NMAIN] = (P_w - D_w + lamda) * volc;
        }
        if (SA && VISC) {   // SA source term (source.f90:835-983)
          const double density = qA[0], tv = qA[5 * PSQ];
          const long long cI = c;   // global index of the cell
          const double wx = rA[(3 * 2 + 1) * PS] - rA[(3 * 1 + 2) * PS], wy = rA[(3 * 0 + 2) * PS] - rA[(3 * 2 + 0) * PS],
                       wz = rA[(3 * 1 + 0) * PS] - rA[(3 * 0 + 1) * PS];
          const double vort = sqrt(((wx * wx) + (wy * wy) + (wz * wz)));
          const double tvx = rA[(3 * 4 + 0) * PS], tvy = rA[(3 * 4 + 1) * PS], tvz = rA[(3 * 4 + 2) * PS];
          const double CD1 = kCb2 * ((tvx * tvx) + (tvy * tvy) + (tvz * tvz));
          const double CD2 = rA[15 * PS];   // grad(rho) . grad(nu-tilde), made by k_gradients (16th staged gradient field)
          const double mu_c = rA[S::OFF_MU * PS];
          const double dist_c = a.geom[(long long)G_DIST * fs + cI];
          const double kd = kKappaSA * dist_c, kd2 = kd * kd;
          const double nu = mu_c / density;
          const double xi = tv / nu;
          const double fv1 = (pow3(xi)) / ((pow3(xi)) + (pow3(kCv1)));
          const double fv2 = 1.0 - xi / (1.0 + (xi * fv1));
          const double scap = fmax(vort + (tv * fv2 / (kd2)), 0.3 * vort);
          const double r = fmin(tv / (scap * kd2), 10.0);
          const double fw = sa_fw(r);
          const double td = tv / dist_c;
          const double D_v = density * kCw1 * fw * (td * td);
          const double P_v = density * kCb1 * scap * tv;
          const double lamda = density * CD1 / kSigmaSA - CD2 * (nu + tv) / kSigmaSA;
          double S_v = (P_v - D_v + lamda) * volc;
          if (RARE && P.trans_bc) {   // add_saBC_source (source.f90:985-1194); its destruction term carries no density (:1181)
            const double u_ = qA[PSQ], v_ = qA[2 * PSQ], w_ = qA[3 * PSQ];
            const double vmag = sqrt(u_ * u_ + v_ * v_ + w_ * w_);
            const double dist2 = dist_c * dist_c;
            const double inv_k2_d2 = 1.0 / ((kKappaSA * kKappaSA) * dist2);
            const double Shat = fmax(vort + tv * fv2 * inv_k2_d2, 1.0e-10);
            const double inv_Shat = 1.0 / Shat;
            const double gBC = gamma_bc(P.re_theta_t, P.nu_cr, tv * fv1, vmag, dist_c, dist2 * vort / nu);
            const double Production = gBC * kCb1 * Shat * tv * volc;
            const double fwb = sa_fw(fmin(tv * inv_Shat * inv_k2_d2, 10.0));
            const double Destruction = (kCw1 * fwb * tv * tv / dist2) * (volc);
            const double lam2 = (density * CD1 / kSigmaSA - CD2 * (nu + tv) / kSigmaSA) * volc;
            S_v = (Production - Destruction + lam2);
          }
          pk[NMAIN] = vort;
          pk[2 * NMAIN] = S_v;
          pk[3 * NMAIN] = mu_c;
          pk[4 * NMAIN] = dist_c;
        }
      }
    }
    // ---- cell work of plane k-1 (low-j halo warp: rows 0..ROWS_JL-1, W_C: the rest) --------------------------------------------
    if (r_hi > r_lo && k - 1 >= kb) {
      const int ic = i0 + lane;
#pragma unroll 1
      for (int r = r_lo; r < r_hi; ++r) {
        const int jc = j0 + r;
        if (ic <= Ly.imx - 1 && jc <= Ly.jmx - 1)
          cell_work<NV, VISC>(P, a, smem, lane, r, ic, jc, k - 1, need_dt, k_active, smem + S::OFF_NRM + (wid == W_C ? 32 : 0) + lane, RARE && P.kkl);
      }
    }
    cp_async_wait_all();
  }
  PT3_FLUSH

  if (a.want_norms) {   // per-CTA partial: warp shuffle inside the two warps that did the cell work, then across them
    bar_all();
    double* sred = smem;   // [NV+1][2]
    const bool cw = wid == W_JL || wid == W_C;
    if (cw) {
      double x[NV + 1];
#pragma unroll
      for (int v = 0; v <= NV; ++v) x[v] = smem[S::OFF_NRM + v * 64 + (wid == W_C ? 32 : 0) + lane];
#pragma unroll
      for (int v = 0; v <= NV; ++v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x[v] += __shfl_down_sync(0xffffffffu, x[v], o);
      }
      if (lane == 0) {
#pragma unroll
        for (int v = 0; v <= NV; ++v) sred[v * 2 + (wid == W_C ? 1 : 0)] = x[v];
      }
    }
    bar_all();
    if (tid <= NV) {
      const double x = sred[tid * 2] + sred[tid * 2 + 1];
      const long long cta = blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z);
      a.red[cta * (NV + 1) + tid] = x;
    }
  }
}

// k planes per CTA: long enough to amortise the two extra iterations, short enough for >= ~4 waves of CTAs
static int pick_kchunk(const Layout& L) {
  const int nk = L.kmx - 1;
  const long long tiles = (long long)((L.imx - 1 + TX - 1) / TX) * ((L.jmx - 1 + TY - 1) / TY);
  int chunk = nk;
  while (chunk > 16 && tiles * ((nk + chunk - 1) / chunk) < 148 * 4) chunk = (chunk + 1) / 2;
  return chunk;
}

template <int NV, int INTERP, int SCHEME, bool VISC, bool RARE>
static int launch_one(Ctx* ctx, KArgs& a) {
  const Layout& L = ctx->P.L;
  a.kchunk = pick_kchunk(L);
  dim3 grid((L.imx - 1 + TX - 1) / TX, (L.jmx - 1 + TY - 1) / TY, (L.kmx - 1 + a.kchunk - 1) / a.kchunk);
  const size_t shm = sizeof(double) * Sm<NV, VISC>::TOTAL;
  static bool attr_set[64] = {false};   // per instantiation and device
  if (!attr_set[ctx->device & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k_sweep3<NV, INTERP, SCHEME, VISC, RARE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    if (e != cudaSuccess) return F3D_ERR_CUDA;
    attr_set[ctx->device & 63] = true;
  }
  if (!ctx->tmaps_ok) return F3D_ERR_CUDA;
  TMaps tm;
  tm.q = (a.q == ctx->tm_q_ptr[0]) ? ctx->tm_q[0] : ctx->tm_q[1];
  if (a.q != ctx->tm_q_ptr[0] && a.q != ctx->tm_q_ptr[1]) return F3D_ERR_ARGUMENT;
  tm.grad = ctx->tm_grad; tm.aux = ctx->tm_aux;
  k_sweep3<NV, INTERP, SCHEME, VISC, RARE><<<grid, NT, shm, ctx->stream>>>(ctx->P, a, tm);
  ctx->launches++;
  return 0;
}

template <int NV, bool VISC, bool RARE>
static int launch_interp(Ctx* ctx, KArgs& a) {
  switch (ctx->P.interpolant) {
    case F3D_INTERP_NONE: return launch_one<NV, F3D_INTERP_NONE, -1, VISC, RARE>(ctx, a);
    case F3D_MUSCL:
      // the headline configuration, in both instantiation sets: with the run-time scheme switch (all six schemes in one kernel) the same
      // case runs a third slower (sst + transition bc, 256^3: 8.9 vs 6.x ms per step)
      if (ctx->P.scheme == F3D_AUSM) return launch_one<NV, F3D_MUSCL, F3D_AUSM, VISC, RARE>(ctx, a);
      return launch_one<NV, F3D_MUSCL, -1, VISC, RARE>(ctx, a);
    case F3D_PPM: return launch_one<NV, F3D_PPM, -1, VISC, RARE>(ctx, a);
    case F3D_WENO:
      if (!RARE && ctx->P.scheme == F3D_AUSMP) return launch_one<NV, F3D_WENO, F3D_AUSMP, VISC, RARE>(ctx, a);   // BASELINE's second synthetic configuration
      return launch_one<NV, F3D_WENO, -1, VISC, RARE>(ctx, a);
    case F3D_WENO_NM: return launch_one<NV, F3D_WENO_NM, -1, VISC, RARE>(ctx, a);
  }
  return F3D_ERR_UNSUPPORTED;
}

template <bool RARE>
static int launch_sweep3_set(Ctx* ctx, KArgs& a) {
  if (ctx->P.sa) return ctx->P.viscous ? launch_interp<6, true, RARE>(ctx, a) : F3D_ERR_UNSUPPORTED;   // sa needs mu_ref /= 0
  if (ctx->P.lctm) {   // eight variables: compiled into the second set only (sweep3_rare.cu)
    if constexpr (RARE) return ctx->P.viscous ? launch_interp<8, true, true>(ctx, a) : F3D_ERR_UNSUPPORTED;
    else return F3D_ERR_UNSUPPORTED;
  }
  if (ctx->P.viscous) return ctx->P.sst ? launch_interp<7, true, RARE>(ctx, a) : launch_interp<5, true, RARE>(ctx, a);
  return launch_interp<5, false, RARE>(ctx, a);
}

}  // namespace g3
}  // namespace f3d
