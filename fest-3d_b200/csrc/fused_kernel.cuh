// Fused residual / time-step / update kernel, generation 4: Green-Gauss gradients, mu / mu_t / F1 and the ghost-gradient rule are
// computed INSIDE the tile pass -- no gradient, viscosity, face-state, flux or residual array exists in HBM.
//
// One CTA (16 warps, 128 registers per thread) owns a 32 x 4 column of cells and marches through a chunk of k planes.  What sits
// where while plane k is being worked on:
//
//   shared memory   q ring, 4 planes (k .. k+3): nv primitive fields + Temp, boxes of 36 x 8 cells (tile + two ring cells), filled
//                   by the TMA engine (cp.async.bulk.tensor.4d, completion on one mbarrier per slot), one plane per iteration, a
//                   whole iteration ahead of its first use
//                   geometry ring, 3 planes: volume + cell centre (36 x 6 boxes, TMA)
//                   record ring, 2 planes (k, k+1): per cell of tile + ring the 3 x n_grad gradient components, mu, mu_t, F1 as
//                   ONE array-of-structures record (pitch 22 doubles) so that a face reads its two records with LDS.128
//                   exchange area of the i / j faces: hi values, then fluxes (same slot), single-buffered
//   tensor memory   (tcgen05.alloc, 128 columns; a warp reaches the 32 lanes of its own quarter): everything that stays with one
//                   cell column -- the k-face flux ring, the carried k-face value, the cell packet (volume, F1, source terms) and
//                   the norm partials.  Row r of the tile is served by warps r (I row), 4+r (J row), 8+r (K row), 12+r (cell
//                   work), all in lane quarter r, so these hand-overs never touch shared memory (LDTM / STTM move 16 doubles per
//                   instruction where shared memory needs 16 LDS.64).
//
// Per plane, two phases separated by CTA barriers:
//   phase 1  all flux warps, one instruction stream: I rows, J rows, K rows + the three halo warps reconstruct, exchange and
//            evaluate ONE face per thread (inviscid + viscous + time-step face terms); the I rows also leave the cell packet
//   phase 2  warps 0..6: Green-Gauss gradients + viscosities of plane k+2 for the 34 x 6 cells of tile + ring from the staged q
//            planes k+1, k+2, k+3 (face metrics straight from global memory / L2), then the ghost-gradient rule of the physical
//            faces inside the tile; warps 12..15: cell work of plane k (residual assembly, source, local time step,
//            point-implicit scaling, RK accumulation, update, norm partials)
//
// Reference pipeline reproduced: src/update.f90:534-545, 228-491; src/face/state/*.f90; src/boundary/boundary_state_reconstruction.f90:
// 93-131; src/face/flux/convective/*.f90, scheme.f90:111-141; src/gradients.f90:405-482 (compute_gradient_G), :486-676 (ghost rule,
// with the mis-indexed face records gathered at set-up); src/viscosity.f90:109-138, 149-163, 215-263, 343-465; src/viscous.f90:
// 144-447, 570-656; src/source.f90:158-270, 467-604, 835-1194; src/time.f90:122-246, 366-531; src/resnorm.f90:171-199.
#pragma once
#include "sweep_common.cuh"

namespace f3d {
namespace g4 {

#ifdef F3D_PHASE_TIMING   // development aid: clock64 totals per warp and bucket (phase 1 work, wait at barrier A, phase 2 work, wait at barrier B)
__device__ unsigned long long g_phase[16 * 4];
#define PT_DECL unsigned long long pt_t = clock64(), pt_acc[4] = {0, 0, 0, 0};
#define PT_MARK(n) { const unsigned long long t_ = clock64(); pt_acc[n] += t_ - pt_t; pt_t = t_; }
#define PT_FLUSH if (lane == 0) { for (int n_ = 0; n_ < 4; ++n_) atomicAdd(&g_phase[wid * 4 + n_], pt_acc[n_]); }
#else
#define PT_DECL
#define PT_MARK(n)
#define PT_FLUSH
#endif

constexpr int TX = 32, TY = 4;
constexpr int NMAIN = TX * TY;
constexpr int NW = 15, NT = 32 * NW;                            // 15 warps x 136 registers = the register file of the SM
constexpr int W_J0 = TY, W_K0 = 2 * TY, W_IH = 3 * TY, W_JH = 3 * TY + 1, W_JL = 3 * TY + 2;
constexpr int W_OBS = W_K0 + 3;                                 // a warp without phase-2 work: observes the plane arrivals of inviscid runs
// cell work of row r (phase 2) by a warp of the row's lane quarter: the three halo warps 12..14 for rows 0..2, J row 3 (warp 7) for row 3
__device__ __forceinline__ int cell_row_of_warp(int wid) { return wid >= W_IH ? wid - W_IH : ((wid == W_J0 + 3) ? 3 : -1); }
constexpr int N_IGRP = 32 * (TY + 1), N_JGRP = 32 * (TY + 2);   // threads on named barriers 1 and 2
constexpr int NGW = 7;                                          // warps 0..6 (I rows, J rows 0..2) do the gradient tasks of phase 2 (204 cells)

// staged q plane: rows j0-2 .. j0+TY+1 of PW = TX+4 cells (i0-2 .. i0+TX+1); slot of cell (col, row) = row*PW + col
constexpr int PW = TX + 4, QROWS = TY + 4, PSQ = PW * QROWS;
// record cells: rows j0-1 .. j0+TY of RW = TX+2 cells (i0-1 .. i0+TX); record index rc = row*RW + col
constexpr int RW = TX + 2, RROWS = TY + 2, NRC = RW * RROWS;
// staged geometry plane (volume, centre x, y, z): rows j0-1 .. j0+TY of GW = TX+4 cells (i0-2 ..); slot = row*GW + col
constexpr int GW = TX + 4, PSG = GW * RROWS;
constexpr int NQS = 4, NGS = 3;                                 // ring depths
// exchange area ([field][slot], slot = face): i faces TY x (TX+1), j faces (TY+1) x TX
constexpr int SLOT_I = TY * (TX + 1), SLOT_J = (TY + 1) * TX, EX = SLOT_I + SLOT_J;   // slot EX: dummy (idle lanes of the i-halo warp)
constexpr int EXP = EX + 2;                                     // field pitch of the exchange area

// tensor-memory map of one cell column, in doubles (column = 2 x index): k-face flux ring, cell packet, norm partials, carried hi
constexpr int T_FK = 0;      // [2][10]: flux + lambda / viscous / turbulent time-step terms of the k face below plane p: half p & 1
constexpr int T_PK = 20;     // [6]: volume; sst: F1, S_k, S_w; sa: vorticity, S_v, mu, dist
constexpr int T_NRM = 26;    // [8]: mass imbalance + n_var squared residual sums
constexpr int T_HI = 34;     // [8]: value at the high k face of the cell of the previous plane
constexpr int T_COLS = 128;

template <int NV, bool VISC>
struct Sm : RecF<NV, VISC> {
  using RecF<NV, VISC>::NG;
  using RecF<NV, VISC>::NMU;
  static constexpr int NF = NV + 3;                         // flux + the three face terms of the time step
  static constexpr int NQF = NV + (VISC ? 1 : 0);           // staged q fields: the primitive variables, then Temp
  static constexpr int QSLOT = NQF * PSQ;
  static constexpr int GSLOT = 4 * PSG;
  static constexpr int NRF = VISC ? 3 * NG + NMU : 0;       // record: gradient component c, direction d at 3c+d, then mu [, mu_t [, F1]]
  static constexpr int RP = (NRF + 1) & ~1;                 // record pitch: even (LDS.128 pairs); 22 / 18 / 14 doubles are all
                                                            // conflict-free for 128-bit accesses of consecutive records
  static constexpr int F_MU = 3 * NG;
  static constexpr int RSLOT = NRC * RP;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_G = OFF_Q + NQS * QSLOT;
  static constexpr int OFF_R = OFF_G + NGS * GSLOT;
  static constexpr int OFF_X = OFF_R + 2 * RSLOT;           // exchange [NF][EXP]
  static constexpr int OFF_MBAR = OFF_X + NF * EXP;         // NQS + NGS mbarriers, then the tensor-memory base address
  static constexpr int OFF_RED = OFF_MBAR + NQS + NGS + 1;  // final norm reduction [NV+1][4]
  static constexpr int TOTAL = OFF_RED + (NV + 1) * 4;
};

__device__ __forceinline__ void prefetch_l2(const double* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- mbarrier / TMA / named barriers -----------------------------------------------------------------------------------------------
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
__device__ __forceinline__ void tma_g2s_4d(void* dst, const CUtensorMap* tm, int x, int y, int z, int f, void* b) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(dst)),
               "l"(tm), "r"(x), "r"(y), "r"(z), "r"(f), "r"(smem_u32(b))
               : "memory");
}
struct TMaps { CUtensorMap q, temp, geo; };
__device__ __forceinline__ void bar_all() { asm volatile("bar.sync 0;" ::: "memory"); }
__device__ __forceinline__ void bar_group(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// ---- tensor memory as per-column scratch: tcgen05.ld / tcgen05.st, shape 32x32b (thread t of the warp <-> lane 32*(warp%4) + t) ----
__device__ __forceinline__ void tm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
#define F3D_LO(x) __double2loint(x)
#define F3D_HI(x) __double2hiint(x)
__device__ __forceinline__ void tm_st1(unsigned a, double v0) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(a), "r"(F3D_LO(v0)), "r"(F3D_HI(v0)) : "memory");
}
__device__ __forceinline__ void tm_st2(unsigned a, double v0, double v1) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(F3D_LO(v0)), "r"(F3D_HI(v0)), "r"(F3D_LO(v1)), "r"(F3D_HI(v1)) : "memory");
}
__device__ __forceinline__ void tm_st4(unsigned a, const double* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(a), "r"(F3D_LO(v[0])), "r"(F3D_HI(v[0])), "r"(F3D_LO(v[1])),
               "r"(F3D_HI(v[1])), "r"(F3D_LO(v[2])), "r"(F3D_HI(v[2])), "r"(F3D_LO(v[3])), "r"(F3D_HI(v[3]))
               : "memory");
}
__device__ __forceinline__ void tm_st8(unsigned a, const double* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(a), "r"(F3D_LO(v[0])),
               "r"(F3D_HI(v[0])), "r"(F3D_LO(v[1])), "r"(F3D_HI(v[1])), "r"(F3D_LO(v[2])), "r"(F3D_HI(v[2])), "r"(F3D_LO(v[3])), "r"(F3D_HI(v[3])), "r"(F3D_LO(v[4])),
               "r"(F3D_HI(v[4])), "r"(F3D_LO(v[5])), "r"(F3D_HI(v[5])), "r"(F3D_LO(v[6])), "r"(F3D_HI(v[6])), "r"(F3D_LO(v[7])), "r"(F3D_HI(v[7]))
               : "memory");
}
__device__ __forceinline__ void tm_ld8(unsigned a, double* v) {   // 8 doubles (16 columns); the caller waits (tm_wait_ld) before use
  unsigned r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                 "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(a)
               : "memory");
  tm_wait_ld();
#pragma unroll
  for (int n = 0; n < 8; ++n) v[n] = __hiloint2double(r[2 * n + 1], r[2 * n]);
}
#define F3D_TM_LD32(r, a)                                                                                                                                  \
  asm volatile(                                                                                                                                           \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, " \
      "%26, %27, %28, %29, %30, %31}, [%32];"                                                                                                           \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), \
        "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),       \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                     \
      : "r"(a)                                                                                                                                            \
      : "memory")
__device__ __forceinline__ void tm_ld16_3(unsigned a0, double* v0, unsigned a1, double* v1, unsigned a2, double* v2) {   // 3 x 16 doubles, one wait
  unsigned r0[32], r1[32], r2[32];
  F3D_TM_LD32(r0, a0);
  F3D_TM_LD32(r1, a1);
  F3D_TM_LD32(r2, a2);
  tm_wait_ld();
#pragma unroll
  for (int n = 0; n < 16; ++n) { v0[n] = __hiloint2double(r0[2 * n + 1], r0[2 * n]); v1[n] = __hiloint2double(r1[2 * n + 1], r1[2 * n]); v2[n] = __hiloint2double(r2[2 * n + 1], r2[2 * n]); }
}
__device__ __forceinline__ void tm_ld16(unsigned a, double* v) {   // 16 doubles (32 columns)
  unsigned r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, "
      "%26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
        "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(a)
      : "memory");
  tm_wait_ld();
#pragma unroll
  for (int n = 0; n < 16; ++n) v[n] = __hiloint2double(r[2 * n + 1], r[2 * n]);
}

// ---- viscous face from array-of-structures records (viscous.f90:209-323, 378-446, 570-656; time.f90:396-421, 479-504) --------------
// Same arithmetic as sweep_common.cuh:viscous_face (sums carried instead of halved averages: the reference's bits); what differs is
// where the operands come from: rl / rh = the two cell records (pitch RP, read as 128-bit pairs), cl / ch = staged volume + centre.
template <int NV, int RP>
__device__ __forceinline__ void viscous_face4(const Params& P, const double* __restrict__ ql_, const double* __restrict__ qh_, const double* __restrict__ rl,
                                              const double* __restrict__ rh, const double* __restrict__ cl, const double* __restrict__ ch, double A, double nx,
                                              double ny, double nz, bool sst_on, bool need_dt, double (&F)[NV], double& vis, double& tur) {
  using R = RecF<NV, true>;
  constexpr bool SST = (NV == 7), SA = (NV == 6), TURB = SST || SA;
  constexpr int NG = R::NG, F_MU = 3 * NG;
  const double dx = ch[PSG] - cl[PSG], dy = ch[2 * PSG] - cl[2 * PSG], dz = ch[3 * PSG] - cl[3 * PSG];
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
  double S[RP];   // l + h of every record field
  double mu_hi = 0.0, mut_hi = 0.0;
#pragma unroll
  for (int p = 0; p < RP / 2; ++p) {
    const double2 a = reinterpret_cast<const double2*>(rl)[p], b = reinterpret_cast<const double2*>(rh)[p];
    S[2 * p] = a.x + b.x; S[2 * p + 1] = a.y + b.y;
    if (2 * p == F_MU) mu_hi = b.x;
    if (2 * p + 1 == F_MU) mu_hi = b.y;
    if (TURB && 2 * p == F_MU + 1) mut_hi = b.x;
    if (TURB && 2 * p + 1 == F_MU + 1) mut_hi = b.y;
  }
  double G[NG][3];   // 2 x the face gradient
#pragma unroll
  for (int c = 0; c < NG; ++c) {
    const double ax = S[3 * c], ay = S[3 * c + 1], az = S[3 * c + 2];
    const double nc = ((2. * del[c]) - (ax * dx + ay * dy + az * dz)) * inv_d;
    G[c][0] = ax + (nc * ex);
    G[c][1] = ay + (nc * ey);
    G[c][2] = az + (nc * ez);
  }
  const double mu_s = S[F_MU];
  const double mut_s = TURB ? S[F_MU + 1] : 0.0;
  const double tmu2 = mu_s + mut_s;
  const double tmu = 0.5 * tmu2, tmuh = 0.25 * tmu2;
  const double div3 = (G[0][0] + G[1][1] + G[2][2]) * (1. / 3.);
  const double Txx = tmu * (G[0][0] - div3), Tyy = tmu * (G[1][1] - div3), Tzz = tmu * (G[2][2] - div3);
  const double Txy = tmuh * (G[1][0] + G[0][1]), Txz = tmuh * (G[2][0] + G[0][2]), Tyz = tmuh * (G[2][1] + G[1][2]);
  const double Kh = 0.25 * ((mu_s * P.inv_Pr + mut_s * P.inv_tPr) * P.gm * P.R_gas * P.inv_gm1);
  const double Qx = Kh * G[3][0], Qy = Kh * G[3][1], Qz = Kh * G[3][2];
  const double A4 = 0.25 * A;
  const double uf = 0.5 * (ql[1] + qh[1]), vf = 0.5 * (ql[2] + qh[2]), wf = 0.5 * (ql[3] + qh[3]);
  F[1] = F[1] - ((Txx * nx + Txy * ny + Txz * nz) * A);
  F[2] = F[2] - ((Txy * nx + Tyy * ny + Tyz * nz) * A);
  F[3] = F[3] - ((Txz * nx + Tyz * ny + Tzz * nz) * A);
  F[4] = F[4] - (A * (((Txx * uf + Txy * vf + Txz * wf + Qx) * nx) + ((Txy * uf + Tyy * vf + Tyz * wf + Qy) * ny) +
                      ((Txz * uf + Tyz * vf + Tzz * wf + Qz) * nz)));
  if (SST && sst_on) {
    const double F1 = 0.5 * S[F_MU + 2];
    const double sk = kSigmaK1 * F1 + kSigmaK2 * (1.0 - F1);
    const double sw = kSigmaW1 * F1 + kSigmaW2 * (1.0 - F1);
    const double rhof = 0.5 * (ql[0] + qh[0]);
    const double tkf = 0.5 * (ql[NV - 2] + qh[NV - 2]);
    const double Tk = -2.0 * rhof * tkf * (1. / 3.);
    const double dk = (A4 * ((mu_s + sk * mut_s) * (G[NG - 2][0] * nx + G[NG - 2][1] * ny + G[NG - 2][2] * nz)));
    const double dw = (A4 * ((mu_s + sw * mut_s) * (G[NG - 1][0] * nx + G[NG - 1][1] * ny + G[NG - 1][2] * nz)));
    F[1] = F[1] - (Tk * nx * A);
    F[2] = F[2] - (Tk * ny * A);
    F[3] = F[3] - (Tk * nz * A);
    F[4] = F[4] - dk;
    F[NV - 2] = F[NV - 2] - dk;
    F[NV - 1] = F[NV - 1] - dw;
  }
  if (SA) {   // viscous.f90:570-656: its "mut_f" is rho_face * nu-tilde_face, not the eddy viscosity; K flux also when kmx == 2
    const double rhof = 0.5 * (ql[0] + qh[0]);
    const double mut_sa = 0.5 * (ql[5] + qh[5]) * rhof;
    F[5] = F[5] - ((0.5 * A) * (((0.5 * mu_s) + mut_sa) * (G[4][0] * nx + G[4][1] * ny + G[4][2] * nz))) * (1.0 / kSigmaSA);
  }
  if (need_dt) {
    const double dn = fabs(((-dx) * nx) + ((-dy) * ny) + ((-dz) * nz));
    const double w = A * rcp64(qh[0] * dn);
    vis = w * mu_hi;
    if (TURB) tur = w * mut_hi;
  }
}

// One face: boundary overrides of the states (boundary_state_reconstruction.f90:124-131), inviscid flux times area (scheme.f90:68-109),
// viscous flux, the face terms of the time step.  f = face index along direction d, m = node count along d.
template <int NV, int SCHEME, bool VISC, int RP>
__device__ __forceinline__ void face_eval4(const Params& P, int d, const double* __restrict__ ql, const double* __restrict__ qh, const double* __restrict__ rl,
                                           const double* __restrict__ rh, const double* __restrict__ cl, const double* __restrict__ ch, double A, double nx,
                                           double ny, double nz, int f, int m, double (&L)[NV], double (&R)[NV], bool flux_on, bool need_dt, double (&F)[NV],
                                           double& lam, double& vis, double& tur) {
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
  if constexpr (VISC) viscous_face4<NV, RP>(P, ql, qh, rl, rh, cl, ch, A, nx, ny, nz, (NV == 7) && flux_on, need_dt, F, vis, tur);
}

// ---- phase 2, gradient task: Green-Gauss gradients (gradients.f90:405-482) + Sutherland / SA / SST viscosities (viscosity.f90) of one
// cell of plane p from the staged q planes p-1 (qm), p (q0), p+1 (qp); sq = the cell's slot in a q plane.  Writes the record.
// n*A of the six faces of a cell per direction component: the only global loads of the gradient task (rows of consecutive cells;
// the same lines come back from L2 when the flux warps ask for them two planes later).  Requested BEFORE the phase barrier, so
// that their latency is spent waiting at the barrier instead of at the head of the task.
struct GradW { double wl[3][3], wh[3][3]; };
__device__ __forceinline__ void gradient_weights(const KArgs& a, const Layout& L, long long c, GradW& w) {
  const long long fs = L.fs;
  const double* __restrict__ gI = a.geom + (long long)G_IA * fs + c;
  const double* __restrict__ gJ = a.geom + (long long)G_JA * fs + c;
  const double* __restrict__ gK = a.geom + (long long)G_KA * fs + c;
  const double AIl = gI[0], AIh = gI[1], AJl = gJ[0], AJh = gJ[L.sj], AKl = gK[0], AKh = gK[L.sk];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    w.wl[0][d] = gI[(1 + d) * fs] * AIl; w.wh[0][d] = gI[(1 + d) * fs + 1] * AIh;
    w.wl[1][d] = gJ[(1 + d) * fs] * AJl; w.wh[1][d] = gJ[(1 + d) * fs + L.sj] * AJh;
    w.wl[2][d] = gK[(1 + d) * fs] * AKl; w.wh[2][d] = gK[(1 + d) * fs + L.sk] * AKh;
  }
}

__device__ __forceinline__ void gradient_prefetch(const KArgs& a, const Layout& L, long long c) {
  const long long fs = L.fs;
  const double* gI = a.geom + (long long)G_IA * fs + c;
#pragma unroll
  for (int f = 0; f < 12; ++f) prefetch_l2(gI + f * fs);   // A, nx, ny, nz of the I, J, K faces of the cell: the neighbours' lines cover the high faces
  prefetch_l2(a.geom + (long long)G_DIST * fs + c);
}

// tanh(x) for x >= 0 through one exponential: (1 - e) / (1 + e), e = exp(-2x).  Absolute error <= ~2e-16 (the blending functions it
// feeds are O(1) factors: the error that matters is absolute), a third of the instructions of the library tanh and no slow path.
__device__ __forceinline__ double tanh_pos(double x) {
  const double e = exp(-2.0 * dmin(x, 20.0));
  return (1.0 - e) * rcp64(1.0 + e);
}

// SST eddy viscosity and F1 of a cell from its gradients (viscosity.f90:215-263, 343-388)
__device__ __forceinline__ void sst_eddy4(const Params& P, double density, double tk, double tw, double mu, double dd, const double (&g)[6][3], double& mut,
                                          double& F1) {
  const double var1 = sqrt(tk) * rcp64(kBstar * tw * dd);
  const double var2 = 500 * (mu * rcp64(density)) * rcp64((dd * dd) * tw);
  const double arg2 = dmax(2 * var1, var2);
  const double Fb = tanh_pos(arg2 * arg2);
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
  const double right = 4 * (density * kSigmaW2 * tk) * rcp64(CD * (dd * dd));
  const double left = dmax(var1, var2);
  const double arg1 = dmin(left, right);
  F1 = tanh_pos((arg1 * arg1) * (arg1 * arg1));
}

template <int NV>
__device__ __forceinline__ void gradient_record(const Params& P, const KArgs& a, const double* __restrict__ qm, const double* __restrict__ q0,
                                                const double* __restrict__ qp, int sq, double vol_c, long long c, const GradW& w,
                                                double* __restrict__ rec) {
  using R = RecF<NV, true>;
  constexpr int NG = R::NG, F_MU = 3 * NG, RP = ((3 * NG + R::NMU) + 1) & ~1;
  const Layout& L = P.L;
  const long long fs = L.fs;
  const double (&wl)[3][3] = w.wl;
  const double (&wh)[3][3] = w.wh;
  const double ivol2 = rcp64(2 * vol_c);
  const bool zgrad = L.kmx > 2;   // gradqp_z = 0 when kmx == 2 (gradients.f90:328-336)
  double g[NG][3];
  double nan_probe = 0.0;
#pragma unroll
  for (int cc = 0; cc < NG; ++cc) {
    const int f = (cc == 3) ? NV : cc + 1;   // u, v, w, Temp (staged behind the primitive variables), then the turbulence variables
    const double v0 = q0[f * PSQ + sq];
    const double sIl = q0[f * PSQ + sq - 1] + v0, sJl = q0[f * PSQ + sq - PW] + v0, sKl = qm[f * PSQ + sq] + v0;
    const double sIh = q0[f * PSQ + sq + 1] + v0, sJh = q0[f * PSQ + sq + PW] + v0, sKh = qp[f * PSQ + sq] + v0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      double r = (-sIl * wl[0][d] - sJl * wl[1][d] - sKl * wl[2][d] + sIh * wh[0][d] + sJh * wh[1][d] + sKh * wh[2][d]) * ivol2;
      if (d == 2) r = zgrad ? r : 0.0;
      nan_probe += r;
      g[cc][d] = r;
    }
  }
  if (isnan(nan_probe)) atomicOr(a.err, F3D_ERR_NAN_GRADIENT);
  const double density = q0[sq], pres = q0[4 * PSQ + sq];
  double mu = P.mu_ref;   // constant viscosity: mu = mu_ref everywhere (viscosity.f90:527)
  if (P.mu_variation == 1) {
    const double T = pres * rcp64(density * P.R_gas);
    const double tr = T / P.T_ref;   // (T/T_ref)**1.5 = tr*sqrt(tr): <= 1 ulp from pow, 8x cheaper
    mu = P.mu_ref * (tr * sqrt(tr)) * ((P.T_ref + P.Sutherland_temp) * rcp64(T + P.Sutherland_temp));
    if (isnan(mu)) atomicOr(a.err, F3D_ERR_NAN_VISCOSITY);
  }
  double mut = 0.0, F1 = 0.0;
  if (NG == 5) {   // Spalart-Allmaras: mu_t = rho*tv*fv1 (viscosity.f90:149-163)
    const double tv = q0[5 * PSQ + sq];
    const double xi = tv * density / mu;
    const double fv1 = (pow3(xi)) / ((pow3(xi)) + (pow3(kCv1)));
    mut = density * tv * fv1;
  }
  if constexpr (NG == 6) sst_eddy4(P, density, q0[5 * PSQ + sq], q0[6 * PSQ + sq], mu, a.geom[(long long)G_DIST * fs + c], g, mut, F1);
  double out[RP];
#pragma unroll
  for (int cc = 0; cc < NG; ++cc) { out[3 * cc] = g[cc][0]; out[3 * cc + 1] = g[cc][1]; out[3 * cc + 2] = g[cc][2]; }
  out[F_MU] = mu;
  if (R::NMU > 1) out[F_MU + 1] = mut;
  if (R::NMU > 2) out[F_MU + 2] = F1;
  if (RP > 3 * NG + R::NMU) out[RP - 1] = 0.0;
#pragma unroll
  for (int p = 0; p < RP / 2; ++p) reinterpret_cast<double2*>(rec)[p] = make_double2(out[2 * p], out[2 * p + 1]);
}

// Ghost-gradient rule + ghost mu_t / F1 of one ghost cell next to a physical face (gradients.f90:638-674, viscosity.f90:165-212, 408-465).
// recI = the finished record of the interior cell behind the face, rec = the ghost cell's own record (overwritten); qI / qG point at
// density of the interior / ghost cell in their staged q planes; fr = the (mis-indexed) face record A, nx, ny, nz; face = 1..6.
template <int NV>
__device__ __forceinline__ void ghost_record(const Params& P, const double* __restrict__ recI, double* __restrict__ rec, const double* __restrict__ qI,
                                             const double* __restrict__ qG, const double* __restrict__ fr, double vol_i, int face, double dist_g) {
  using R = RecF<NV, true>;
  constexpr int NG = R::NG, F_MU = 3 * NG;
  const bool lo = (face % 2) == 1;
  const double A = fr[0], nx = fr[1], ny = fr[2], nz = fr[3];
  const double c_x = A * nx / vol_i, c_y = A * ny / vol_i, c_z = A * nz / vol_i;
  const double sig = lo ? 1.0 : -1.0;
  const int id = P.bc_id[face - 1];
  const double ft = P.fixed[F3D_FIX_WALL_TEMP][face - 1];
#pragma unroll
  for (int cc = 0; cc < NG; ++cc) {
    // slot cc holds variable cc+2 of qp(2:n_var): u,v,w,p,[k,omega]; slot 4 (cc == 3) is then overwritten with T
    const int f = (cc == 3) ? NV : cc + 1;
    const double vI = qI[f * PSQ], vG = qG[f * PSQ];
    const double gIx = recI[3 * cc], gIy = recI[3 * cc + 1], gIz = recI[3 * cc + 2];
    double gx = sig * (vI - vG) * c_x, gy = sig * (vI - vG) * c_y, gz = sig * (vI - vG) * c_z;
    if (cc == 3 && id == -5 && (ft < 1. && ft >= 0.)) { gx = -gIx; gy = -gIy; gz = -gIz; }   // adiabatic wall
    const double dot = (gIx * nx) + (gIy * ny) + (gIz * nz);
    rec[3 * cc] = gx + (gIx - dot * nx);
    rec[3 * cc + 1] = gy + (gIy - dot * ny);
    rec[3 * cc + 2] = gz + (gIz - dot * nz);
  }
  const bool copyish = id == -1 || id == -2 || id == -3 || id == -4 || id == -6 || id == -7 || id == -8 || id == -9;
  if (R::NMU >= 2) {
    if (id == -5) rec[F_MU + 1] = -recI[F_MU + 1];
    else if (copyish) rec[F_MU + 1] = recI[F_MU + 1];
  }
  if (R::NMU >= 3) {
    if (id == -5 || copyish) rec[F_MU + 2] = recI[F_MU + 2];
  }
  if constexpr (NG == 6) {
    // Where the reference does not copy (periodic interfaces -10, total pressure -11) the ghost mu_t / F1 follow from the ghost cell's own
    // state and its rule-made gradients: apply_gradient_bc runs before calculate_viscosity (update.f90:534-541)
    if (id != -5 && !copyish) {
      double g[6][3];
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) { g[cc][0] = rec[3 * cc]; g[cc][1] = rec[3 * cc + 1]; g[cc][2] = rec[3 * cc + 2]; }
      double mut, F1;
      sst_eddy4(P, qG[0], qG[5 * PSQ], qG[6 * PSQ], rec[F_MU], dist_g, g, mut, F1);
      rec[F_MU + 1] = mut; rec[F_MU + 2] = F1;
    }
  }
}

// ---- phase 2, cell work of one cell (row r of the tile, plane kc): residual assembly from the six face fluxes, source, local time
// step, point-implicit k/omega scaling, RK accumulation, conservative update, norm partials.  i/j fluxes from the exchange area,
// everything that belongs to the cell column from tensor memory: Flo / Fhi = the k faces below / above, pk = cell packet, nrm.
template <int NV, bool VISC>
__device__ __forceinline__ void cell_work4(const Params& P, const KArgs& a, const double* __restrict__ xF, int tx, int r, int i, int j, int kc, bool need_dt,
                                           bool k_active, const double* Flo, const double* Fhi, const double* pk, double* nrm) {
  constexpr bool SST = (NV == 7), SA = (NV == 6), TURB = SST || SA;
  const Layout& Ly = P.L;
  const long long fs = Ly.fs;
  const long long cc = Ly.idx(i, j, kc);
  const int sl0 = r * (TX + 1) + tx, sh0 = sl0 + 1, sl1 = SLOT_I + r * TX + tx, sh1 = sl1 + TX;
  double qc[NV];   // the state of the cell: from global memory (L2: it was staged a few planes ago)
#pragma unroll
  for (int v = 0; v < NV; ++v) qc[v] = a.q[v * fs + cc];
  const double volc = pk[0];
  double res[NV];
  double merr = 0.0;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const double Fl0 = xF[v * EXP + sl0], Fh0 = xF[v * EXP + sh0], Fl1 = xF[v * EXP + sl1], Fh1 = xF[v * EXP + sh1];
    double rr = 0.0;
    rr = rr + (Fh0 - Fl0);   // scheme.f90:133-135
    rr = rr + (Fh1 - Fl1);
    if (k_active) rr = rr + (Fhi[v] - Flo[v]);
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
    res[5] = res[5] - pk[2];
    res[6] = res[6] - pk[3];
  }
  if (SA && VISC) res[5] = res[5] - pk[2];

  double dtc = 0.0;
  if (need_dt) {
    if (P.time_stepping == 1 && P.global_time_step > 0) {
      dtc = P.global_time_step;
    } else {
      const double* lamv = xF + NV * EXP;
      const double lmxsum = lamv[sl0] + lamv[sl1] + Flo[NV] + lamv[sh0] + lamv[sh1] + Fhi[NV];
      dtc = rcp64(lmxsum);
      dtc = dtc * volc * P.CFL;
      if (VISC) {
        const double* visv = xF + (NV + 1) * EXP;
        double s = visv[sl0] + visv[sl1] + Flo[NV + 1] + visv[sh0] + visv[sh1] + Fhi[NV + 1];
        s = P.gm * s * P.inv_Pr;
        s = 2. * rcp64(s + (2. * P.CFL * volc * rcp64(dtc)));
        dtc = P.CFL * (s * volc);
        if (TURB) {
          const double* turv = xF + (NV + 2) * EXP;
          double tt = turv[sl0] + turv[sl1] + Flo[NV + 2] + turv[sh0] + turv[sh1] + Fhi[NV + 2];
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
    } else {
      u1[0] = qc[0];
#pragma unroll
      for (int v = 1; v < NV; ++v) u1[v] = qc[v] * u1[0];
    }
    u1[4] = (u1[4] * P.inv_gm1 + 0.5 * (u1[1] * u1[1] + u1[2] * u1[2] + u1[3] * u1[3])) * rcp64(u1[0]) + 0.;
    if (SST) {
      const double F1 = VISC ? pk[1] : 0.0;
      const double beta = kBeta1 * F1 + (1. - F1) * kBeta2;
      R[5] = R[5] * rcp64(1 + (beta * qc[6] * dtc));
      R[6] = R[6] * rcp64(1 + (2 * beta * qc[6] * dtc));
    }
    if (SA && VISC) {   // update.f90:405-420: u1(6) is rho*tv here, used where the model has tv -- reproduced
      const double vort = pk[1], mu_c = pk[3], dist_c = pk[4];
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
      }
      if (SA) a.qnew[5 * fs + cc] = fmax(u2[5], 1.e-12);   // update.f90:474-475
    }
  }
  if (a.want_norms) {   // resnorm.f90:187-198
    nrm[0] += merr;
#pragma unroll
    for (int v = 0; v < NV; ++v) nrm[1 + v] += res[v] * res[v];
  }
}


// RARE gates the code of the seldom-used options (pressure-based switching, transition = bc): compiled into a second set of
// instantiations (fused_rare.cu) so that the register-tight common path does not carry them.
static_assert(TX == kG3TX && TY == kG3TY, "tensor-map boxes are encoded for this tile (api.cu)");
template <int NV, int INTERP, int SCHEME, bool VISC, bool RARE>
__global__ void __launch_bounds__(NT, 1) k_fused(const Params P, const KArgs a, const __grid_constant__ TMaps tm) {
  using S = Sm<NV, VISC>;
  constexpr bool SST = (NV == 7), SA = (NV == 6), TURB = SST || SA;
  constexpr bool SMQ = (INTERP == F3D_MUSCL || INTERP == F3D_INTERP_NONE);   // 3-point stencils read the staged planes
  constexpr int NF = S::NF, RP = S::RP, F_MU = S::F_MU;
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

  // ---- set-up: mbarriers of the two TMA rings, tensor memory ----------------------------------------------------------------------
  unsigned long long* const mbar = reinterpret_cast<unsigned long long*>(smem + S::OFF_MBAR);
  unsigned* const tm_slot = reinterpret_cast<unsigned*>(smem + S::OFF_MBAR + NQS + NGS);
  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < NQS + NGS; ++b) mbar_init(&mbar[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wid == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tm_slot)), "n"(T_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tm_fence_before();
  __syncthreads();
  tm_fence_after();
  const unsigned tbase = *reinterpret_cast<volatile unsigned*>(tm_slot) + ((unsigned)(32 * (wid & 3)) << 16);   // this warp's lane quarter
  if (cell_row_of_warp(wid) >= 0) {   // the norm partials start at zero (cleared by the warp that accumulates them: no cross-warp ordering needed)
    const double z[8] = {0., 0., 0., 0., 0., 0., 0., 0.};
    tm_st8(tbase + 2 * T_NRM, z);
    tm_wait_st();
  }

  // ---- the two TMA rings.  q plane p sits in slot (p - (kb-2)) & 3, geometry plane p in slot (p - (kb-1)) % 3; tensor coordinates
  // x = i + 15, y = j + 2, z = k + 2 (ctx.hpp:Layout); cells outside the arrays are zero-filled, so the byte counts are those of the
  // full boxes.
  constexpr unsigned q_bytes = 8u * S::QSLOT, g_bytes = 8u * S::GSLOT;
  const int p0 = kb - 2, g0 = kb - 1;
  auto q_off = [&](int p) { return S::OFF_Q + ((p - p0) & (NQS - 1)) * S::QSLOT; };
  auto g_off = [&](int p) { return S::OFF_G + ((p - g0) % NGS) * S::GSLOT; };
  auto r_off = [&](int p) { return S::OFF_R + (p & 1) * S::RSLOT; };
  auto issue_q = [&](int p) {
    const int sl = (p - p0) & (NQS - 1);
    mbar_expect_tx(&mbar[sl], q_bytes);
    tma_g2s_4d(smem + S::OFF_Q + sl * S::QSLOT, &tm.q, i0 + 13, j0, p + 2, 0, &mbar[sl]);
    if (VISC) tma_g2s_4d(smem + S::OFF_Q + sl * S::QSLOT + NV * PSQ, &tm.temp, i0 + 13, j0, p + 2, 0, &mbar[sl]);
  };
  auto issue_g = [&](int p) {
    const int sl = (p - g0) % NGS;
    mbar_expect_tx(&mbar[NQS + sl], g_bytes);
    tma_g2s_4d(smem + S::OFF_G + sl * S::GSLOT, &tm.geo, i0 + 13, j0 + 1, p + 2, 0, &mbar[NQS + sl]);
  };
  auto wait_q = [&](int p) { mbar_wait(&mbar[(p - p0) & (NQS - 1)], (unsigned)(((p - p0) >> 2) & 1)); };
  auto wait_g = [&](int p) { mbar_wait(&mbar[NQS + (p - g0) % NGS], (unsigned)(((p - g0) / NGS) & 1)); };

  if (tid == NT - 1) {   // the first planes; every later one is issued at the start of phase 2, a whole iteration ahead of its use
    issue_q(kb - 2); issue_q(kb - 1); issue_q(kb);
    issue_g(kb - 1);
  }

  // ---- role of this thread in phase 1 (all flux warps run ONE instruction stream, parameterised by these values).  With the staged
  // 3-point stencils (SMQ) the lanes outside the block (ragged last tiles) simply compute on the zero-filled planes: their results
  // land in slots nobody consumes, and the hot path has no divergent domain test.  The 5-point interpolants read global memory and
  // keep the tests.
  // The values are packed into three words and unpacked again at the top of every plane (an empty asm keeps the compiler from
  // hoisting the unpacked forms out of the loop: carried as ~15 separate registers they were spilled, and every plane began by
  // waiting for their reloads -- 17 % of the stall samples of phase 1, profiles/r02_g4_summary.md).
  unsigned rw_ij, rw_sx, rw_fl;
  {
    int i, j, s0, d, row;
    bool rec, fac, wr_hi, irow, krow;
    int exw, exr;                     // I/J: exchange slots: where the hi value goes; where L is read and the flux written
    int rc = 0, sg = 0;
    wr_hi = true; irow = false; krow = false;
    rec = fac = false; d = 0; i = i0 + lane; j = j0; s0 = 2 * PW + 2; rc = RW + 1; sg = GW + 2; exw = exr = EX;
    if (wid < TY) {                   // I row
      const int tx = lane, ty = wid;
      d = 0; i = i0 + tx; j = j0 + ty; irow = true;
      rec = fac = SMQ || ((j <= Ly.jmx - 1) && (i <= Ly.imx));
      s0 = (ty + 2) * PW + tx + 2; rc = (ty + 1) * RW + tx + 1; sg = (ty + 1) * GW + tx + 2;
      exw = ty * (TX + 1) + tx + 1; exr = ty * (TX + 1) + tx;
    } else if (wid < 2 * TY) {        // J row
      const int tx = lane, ty = wid - TY;
      d = 1; i = i0 + tx; j = j0 + ty;
      rec = fac = SMQ || ((i <= Ly.imx - 1) && (j <= Ly.jmx));
      s0 = (ty + 2) * PW + tx + 2; rc = (ty + 1) * RW + tx + 1; sg = (ty + 1) * GW + tx + 2;
      exw = SLOT_I + (ty + 1) * TX + tx; exr = SLOT_I + ty * TX + tx;
    } else if (wid < 3 * TY) {        // K row: the column of its cell
      const int tx = lane, ty = wid - 2 * TY;
      d = 2; i = i0 + tx; j = j0 + ty; krow = true;
      rec = fac = k_active && (SMQ || ((i <= Ly.imx - 1) && (j <= Ly.jmx - 1)));
      s0 = (ty + 2) * PW + tx + 2; rc = (ty + 1) * RW + tx + 1; sg = (ty + 1) * GW + tx + 2;
    } else if (wid == W_IH) {         // the two i columns next to the tile: lanes 0..TY-1 low side, TY..2TY-1 high side; the other
      const int r = lane % TY, side = (lane / TY) & 1;   // lanes repeat them into a dummy exchange slot
      d = 0; i = (side == 0) ? i0 - 1 : i0 + TX; j = j0 + r;
      rec = SMQ || ((lane < 2 * TY) && (j <= Ly.jmx - 1) && (i <= Ly.imx));
      fac = rec && side == 1; wr_hi = side == 0;
      s0 = (r + 2) * PW + (side == 0 ? 1 : TX + 2); rc = (r + 1) * RW + (side == 0 ? 0 : TX + 1); sg = (r + 1) * GW + (side == 0 ? 1 : TX + 2);
      if (lane < 2 * TY) { exw = r * (TX + 1) + (side == 0 ? 0 : TX); exr = r * (TX + 1) + TX; }
    } else {                          // high (W_JH) and low (W_JL) j rows next to the tile
      const bool high = wid == W_JH;
      d = 1; i = i0 + lane; j = high ? j0 + TY : j0 - 1;
      rec = SMQ || ((i <= Ly.imx - 1) && (j <= Ly.jmx));
      fac = rec && high; wr_hi = !high;
      s0 = (high ? TY + 2 : 1) * PW + lane + 2; rc = (high ? TY + 1 : 0) * RW + lane + 1; sg = (high ? TY + 1 : 0) * GW + lane + 2;
      exw = SLOT_I + (high ? TY * TX : 0) + lane; exr = SLOT_I + TY * TX + lane;
    }
    if (i > Ly.imx + 1) i = Ly.imx + 1;
    if (j > Ly.jmx + 1) j = Ly.jmx + 1;
    (void)rc; (void)sg;
    row = s0 / PW;               // rc = s0 - 2 row - (PW - 1), sg = s0 - PW
    rw_ij = (unsigned)i | ((unsigned)j << 16);
    rw_sx = (unsigned)s0 | ((unsigned)exw << 10) | ((unsigned)exr << 20);
    rw_fl = (rec ? 1u : 0u) | (fac ? 2u : 0u) | (wr_hi ? 4u : 0u) | (irow ? 8u : 0u) | (krow ? 16u : 0u) | ((unsigned)d << 5) | ((unsigned)row << 8);
  }
  const bool krow = (wid >= W_K0) && (wid < W_IH), irow = wid < TY;   // warp-uniform forms, for the control flow outside the flux code

  // ---- phase 2, gradient task of the 34 x 6 cells of tile + ring of plane p (warps 0..NGW-1), in two steps around named barrier 3 ------
  const int gt = tid;                             // gradient cell of this thread
  unsigned rw_g;                                  // row | col << 8 | valid << 16
  {
    const int g_row_ = gt / RW, g_col_ = gt - g_row_ * RW;
    const bool v_ = VISC && gt < NRC && (i0 - 1 + g_col_) <= Ly.imx && (j0 - 1 + g_row_) <= Ly.jmx;
    rw_g = (unsigned)g_row_ | ((unsigned)g_col_ << 8) | (v_ ? 0x10000u : 0u);
  }
  int g_row, g_col, gi, gj, g_sq, g_sg;
  bool g_valid;
  auto unpack_g = [&]() {
    unsigned w_ = rw_g;
    asm volatile("" : "+r"(w_));
    g_row = w_ & 0xff; g_col = (w_ >> 8) & 0xff; g_valid = (w_ >> 16) != 0;
    gi = i0 - 1 + g_col; gj = j0 - 1 + g_row;
    g_sq = (g_row + 1) * PW + g_col + 1; g_sg = g_row * GW + g_col + 1;
  };
  auto grad_gauss = [&](int p, const GradW& gw) {
    if constexpr (VISC) {
      if (g_valid)
        gradient_record<NV>(P, a, smem + q_off(p - 1), smem + q_off(p), smem + q_off(p + 1), g_sq, smem[g_off(p) + g_sg], Ly.idx(gi, gj, p), gw,
                            smem + r_off(p) + gt * RP);
    }
  };
  // ghost-gradient rule of the physical faces the tile touches.  In-plane faces (imin .. jmax) and kmax belong to the ghost cell of
  // plane p itself; the kmin rule needs the records of plane 1, so the cell of plane 0 is re-done by the thread that has just
  // finished the same column's cell of plane 1.
  auto grad_ghost = [&](int p) {
    if constexpr (VISC) {
      if (!g_valid) return;
      const bool iin = gi >= 1 && gi <= Ly.imx - 1, jin = gj >= 1 && gj <= Ly.jmx - 1, kin = p >= 1 && p <= Ly.kmx - 1;
      int face = 0, dt_ = 0, pg = p, pi = p;   // ghost cell: record gt of plane pg; interior cell: record gt + dt_ of plane pi
      if (kin && jin) { if (gi == 0) { face = 1; dt_ = 1; } else if (gi == Ly.imx) { face = 2; dt_ = -1; } }
      if (kin && iin) { if (gj == 0) { face = 3; dt_ = RW; } else if (gj == Ly.jmx) { face = 4; dt_ = -RW; } }
      if (iin && jin) { if (p == 1) { face = 5; pg = 0; } else if (p == Ly.kmx) { face = 6; pi = p - 1; } }
      if (face == 0 || P.bc_id[face - 1] >= 0) return;   // "if (bc%imin_id < 0)" -- includes -10
      const int ax = (face - 1) / 2;
      const int ia = (ax == 0) ? gj : gi, ib = (ax == 2) ? gj : pi;          // transverse indices of the face, reference axis order
      const int na = ((ax == 0) ? Ly.jmx : Ly.imx) - 1;
      const double* fr = a.gbc + a.gbc_off[face - 1] + 4 * ((long long)(ib - 1) * na + (ia - 1));
      const int dsq = (dt_ == 1 || dt_ == -1) ? dt_ : ((dt_ == RW) ? PW : ((dt_ == -RW) ? -PW : 0));
      const int dsg = (dt_ == 1 || dt_ == -1) ? dt_ : ((dt_ == RW) ? GW : ((dt_ == -RW) ? -GW : 0));
      const double* qG = smem + q_off(pg) + g_sq;
      const double* qI = smem + q_off(pi) + g_sq + dsq;
      const double vol_i = smem[g_off(pi) + g_sg + dsg];
      const int id_ = P.bc_id[face - 1];
      const double dist_g = (NV == 7 && (id_ == -10 || id_ == -11)) ? a.geom[(long long)G_DIST * Ly.fs + Ly.idx(gi, gj, pg)] : 0.0;
      ghost_record<NV>(P, smem + r_off(pi) + (gt + dt_) * RP, smem + r_off(pg) + gt * RP, qI, qG, fr, vol_i, face, dist_g);
    }
  };

  // The march starts two iterations early: iteration kb-3 only computes the records of plane kb-1, iteration kb-2 those of plane kb
  // and the K rows' value at the high face of cell kb-1 (no face yet); from kb-1 on the K rows evaluate the face between planes k
  // and k+1, from kb on everything runs.  One instance of every piece of code, no separate prologue.
  PT_DECL
  for (int k = kb - 3; k <= ke - 1; ++k) {
    // =========================== phase 1: one reconstruction and one face per thread ====================================================
    const bool active = krow ? k >= kb - 2 : k >= kb;
    if (active) {
      unsigned w_ij = rw_ij, w_sx = rw_sx, w_fl = rw_fl;
      asm volatile("" : "+r"(w_ij), "+r"(w_sx), "+r"(w_fl));
      const int i = w_ij & 0xffff, j = w_ij >> 16;
      const int s0 = w_sx & 0x3ff, exw = (w_sx >> 10) & 0x3ff, exr = w_sx >> 20;
      const bool rec = w_fl & 1, fac = w_fl & 2, wr_hi = w_fl & 4;
      const int d = (w_fl >> 5) & 3, row = w_fl >> 8;
      const int rc = s0 - 2 * row - (PW - 1), sg = s0 - PW;
      const int pos = (d == 0) ? i : j, mx = (d == 0) ? Ly.imx : ((d == 1) ? Ly.jmx : Ly.kmx);
      const int dq = (d == 0) ? 1 : PW, dr = (d == 0) ? 1 : RW, dg = (d == 0) ? 1 : GW;   // low neighbour along d (I / J warps)
      // I/J: cell (i,j,k), face below it along d.  K: cell (i,j,k+1), face between planes k and k+1.  The planes this phase reads
      // landed long ago: the threads of phase 2 waited on their mbarriers, and two CTA barriers lie in between.
      const bool fac_k = fac && k >= kb - 1;   // no face in the priming iteration of the K rows
      const int oqA = q_off(k), oqB = q_off(k + 1);
      const int o_0 = (krow ? oqB : oqA) + s0;                                   // the cell (q field 0)
      const int o_m = krow ? oqA + s0 : o_0 - dq;                                // low stencil neighbour = the cell below the face
      const int o_p = krow ? q_off(k + 2) + s0 : o_0 + dq;                       // high stencil neighbour
      const int r_h = r_off(krow ? k + 1 : k) + rc * RP;                         // records of the two cells of the face
      const int r_l = krow ? r_off(k) + rc * RP : r_h - dr * RP;
      const int g_h = g_off(krow ? k + 1 : k) + sg;                              // their volume / centre
      const int g_l = krow ? g_off(k) + sg : g_h - dg;
      const int cpos = krow ? k + 1 : pos;                                       // index of the cell and of the face along d
      const long long c = Ly.idx(i, j, k);
      const long long cg = krow ? c + Ly.sk : c;                                 // global index of the cell / face
      double gA_ = 0.0, gnx = 0.0, gny = 0.0, gnz = 0.0;   // face metrics, requested before the reconstruction
      if (fac_k) {
        const double* __restrict__ gp = a.geom + (long long)(G_IA + 4 * d) * fs + cg;
        gA_ = gp[0]; gnx = gp[fs]; gny = gp[2 * fs]; gnz = gp[3 * fs];
      }
      double lo[NV], hv[8];
#pragma unroll
      for (int v = 0; v < 8; ++v) hv[v] = 0.0;
      if (rec) {
        double hv_[NV];
        if (SMQ) {
          double qm[NV], q0[NV], qp[NV];
#pragma unroll
          for (int v = 0; v < NV; ++v) { qm[v] = smem[o_m + v * PSQ]; q0[v] = smem[o_0 + v * PSQ]; qp[v] = smem[o_p + v * PSQ]; }
          double p_far = 0.0;   // pressure-based switching at the two ghost positions reads the pressure two cells inwards
          if (RARE && INTERP == F3D_MUSCL && P.pb_switch[d] && (cpos == 0 || cpos == mx)) {
            const int two = (cpos == 0) ? 2 : -2;
            p_far = krow ? q[4 * fs + cg + two * Ly.sk] : smem[o_0 + 4 * PSQ + two * dq];
          }
          recon3<NV, INTERP, RARE>(P, qm, q0, qp, cpos, mx, d, hv_, lo, p_far);
        } else {
          line_cell_values<NV, INTERP, RARE>(P, q, vol, cg, (d == 0) ? 1 : ((d == 1) ? Ly.sj : Ly.sk), cpos, mx, d, hv_, lo);
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) hv[v] = hv_[v];
        if (!krow && wr_hi) {
#pragma unroll
          for (int v = 0; v < NV; ++v) smem[S::OFF_X + v * EXP + exw] = hv[v];
        }
      }
      double L[8];
      if (krow) {   // warp-uniform: L = value at the high face of cell k (left by the previous iteration), then this iteration's one
        tm_ld8(tbase + 2 * T_HI, L);
        tm_st8(tbase + 2 * T_HI, hv);
      } else {
        bar_group(1 + d, (d == 0) ? N_IGRP : N_JGRP);
      }
      double F[10];
#pragma unroll
      for (int v = 0; v < 10; ++v) F[v] = 0.0;
      if (fac_k) {
        double Lf[NV], Ff[NV], lam = 0.0, vis = 0.0, tur = 0.0;
        if (krow) {
#pragma unroll
          for (int v = 0; v < NV; ++v) Lf[v] = L[v];
        } else {
#pragma unroll
          for (int v = 0; v < NV; ++v) Lf[v] = smem[S::OFF_X + v * EXP + exr];
        }
        face_eval4<NV, SCHEME, VISC, RP>(P, d, smem + o_m, smem + o_0, smem + r_l, smem + r_h, smem + g_l, smem + g_h, gA_, gnx, gny, gnz, cpos, mx, Lf, lo,
                                         krow ? flux_on_k : true, need_dt, Ff, lam, vis, tur);
        if (krow) {
#pragma unroll
          for (int v = 0; v < NV; ++v) F[v] = Ff[v];
          F[NV] = lam; F[NV + 1] = vis; F[NV + 2] = tur;
        } else {
#pragma unroll
          for (int v = 0; v < NV; ++v) smem[S::OFF_X + v * EXP + exr] = Ff[v];
          if (need_dt) {
            smem[S::OFF_X + NV * EXP + exr] = lam;
            if (VISC) smem[S::OFF_X + (NV + 1) * EXP + exr] = vis;
            if (VISC && TURB) smem[S::OFF_X + (NV + 2) * EXP + exr] = tur;
          }
        }
      }
      if (krow) {   // flux of the k face below plane k+1 into its half of the ring
        const unsigned ta = tbase + 2 * (T_FK + 10 * ((k + 1) & 1));
        tm_st8(ta, F);
        if (NF > 8) tm_st2(ta + 16, F[8], F[9]);
      }
      if (irow) {   // I rows: the cell packet of the own cell for the cell work of phase 2
        double pkv[6] = {0., 0., 0., 0., 0., 0.};
        if (rec && (SMQ || i <= Ly.imx - 1)) {
          const double* const rA = smem + r_h;   // own record
          const double* const qA = smem + o_0;
          const double volc = smem[g_h];
          pkv[0] = volc;
          if constexpr (SST && VISC) {   // SST source terms (source.f90:214-268)
            double g[6][3];
#pragma unroll
            for (int cc = 0; cc < 6; ++cc) {
              if (cc == 3) continue;
              g[cc][0] = rA[3 * cc + 0]; g[cc][1] = rA[3 * cc + 1]; g[cc][2] = rA[3 * cc + 2];
            }
            const double mut = rA[F_MU + 1];
            const double F1c = rA[F_MU + 2];
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
              const double mu_c = rA[F_MU];
              const double re_v = density * dist_c * dist_c * vort / mu_c;
              P_k = gamma_bc(P.re_theta_t, P.nu_cr, mut / density, vmag, dist_c, re_v) * P_k;
            }
            pkv[1] = F1c;
            pkv[2] = (P_k - D_k) * volc;
            pkv[3] = (P_w - D_w + lamda) * volc;
          }
          if constexpr (SA && VISC) {   // SA source term (source.f90:835-983); the density gradient is built in place from the six neighbours
            const double density = qA[0], tv = qA[5 * PSQ];
            const long long cI = c;   // global index of the cell
            const double rho_km = q[cI - Ly.sk], rho_kp = q[cI + Ly.sk];   // the k neighbours: global memory (L2)
            const double RhoFace[6] = {qA[-1] + density, qA[-PW] + density, rho_km + density, qA[1] + density, qA[PW] + density, rho_kp + density};
            const double* __restrict__ gI = a.geom + (long long)G_IA * fs;
            const double* __restrict__ gJ = a.geom + (long long)G_JA * fs;
            const double* __restrict__ gK = a.geom + (long long)G_KA * fs;
            const long long cf[6] = {cI, cI, cI, cI + 1, cI + Ly.sj, cI + Ly.sk};
            double gradrho[3];
#pragma unroll
            for (int dd = 0; dd < 3; ++dd) {
              // KEPT DEFECT: the normal of the low K face is (nx,nx,nx) (source.f90:901)
              const double n0 = gI[(1 + dd) * fs + cf[0]], n1 = gJ[(1 + dd) * fs + cf[1]], n2 = gK[fs + cf[2]];
              const double n3 = gI[(1 + dd) * fs + cf[3]], n4 = gJ[(1 + dd) * fs + cf[4]], n5 = gK[(1 + dd) * fs + cf[5]];
              gradrho[dd] = (-(RhoFace[0]) * n0 * gI[cf[0]] - (RhoFace[1]) * n1 * gJ[cf[1]] - (RhoFace[2]) * n2 * gK[cf[2]] +
                             (RhoFace[3]) * n3 * gI[cf[3]] + (RhoFace[4]) * n4 * gJ[cf[4]] + (RhoFace[5]) * n5 * gK[cf[5]]) / (2.0 * volc);
            }
            const double wx = rA[3 * 2 + 1] - rA[3 * 1 + 2], wy = rA[3 * 0 + 2] - rA[3 * 2 + 0], wz = rA[3 * 1 + 0] - rA[3 * 0 + 1];
            const double vort = sqrt(((wx * wx) + (wy * wy) + (wz * wz)));
            const double tvx = rA[3 * 4 + 0], tvy = rA[3 * 4 + 1], tvz = rA[3 * 4 + 2];
            const double CD1 = kCb2 * ((tvx * tvx) + (tvy * tvy) + (tvz * tvz));
            const double CD2 = ((gradrho[0] * tvx) + (gradrho[1] * tvy) + (gradrho[2] * tvz));
            const double mu_c = rA[F_MU];
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
            pkv[1] = vort;
            pkv[2] = S_v;
            pkv[3] = mu_c;
            pkv[4] = dist_c;
          }
        }
        tm_st4(tbase + 2 * T_PK, pkv);
        if (SA) tm_st1(tbase + 2 * (T_PK + 4), pkv[4]);
      }
    }
    const bool grad_now = VISC && wid < NGW && k + 2 <= ke;
    tm_wait_st();
    tm_fence_before();
    PT_MARK(0)
    bar_all();   // ---- the fluxes, packets and k faces of plane k are complete; nobody reads q plane k / geometry plane k any more
    tm_fence_after();
    PT_MARK(1)

    // =========================== phase 2: records of plane k+2 | cell work of plane k ===================================================
    if (tid == NT - 1) {   // into the slots of q plane k and geometry plane k
      if (k + 4 <= ke + 1) issue_q(k + 4);
      if (k + 3 <= ke) issue_g(k + 3);
    }
    if (grad_now) {
      if constexpr (VISC) {
        unpack_g();
        GradW gw;
        if (g_valid) gradient_weights(a, Ly, Ly.idx(gi, gj, k + 2), gw);   // L2 hits: prefetched by the previous iteration's task
        if (k == kb - 3) { wait_q(k + 1); wait_q(k + 2); }
        wait_q(k + 3); wait_g(k + 2);
        grad_gauss(k + 2, gw);
        bar_group(3, 32 * NGW);
        grad_ghost(k + 2);
        if (g_valid && k + 3 <= ke) gradient_prefetch(a, Ly, Ly.idx(gi, gj, k + 3));   // next plane's face metrics on their way to L2
      }
    } else if (wid == W_OBS && k + 2 <= ke + 1) {
      // inviscid runs have no gradient task: one warp observes the arrival of the planes for everybody (the barrier below orders it)
      if (k == kb - 3) { wait_q(k + 1); wait_q(k + 2); }
      if (k + 3 <= ke + 1) wait_q(k + 3);
      if (k + 2 <= ke) wait_g(k + 2);
    }
    const int crow = cell_row_of_warp(wid);
    if (crow >= 0 && k >= kb) {   // cell work of row crow (the warp shares the lane quarter of the row's I / J / K warps)
      const int r = crow, ic = i0 + lane, jc = j0 + r;
      double Flo[16], Fhi[16], pn[16];
      tm_ld16_3(tbase + 2 * (T_FK + 10 * (k & 1)), Flo,          // k face below plane k
                tbase + 2 * (T_FK + 10 * ((k + 1) & 1)), Fhi,    // and above it
                tbase + 2 * T_PK, pn);                           // cell packet [0..5], norm partials [6..13]
      if (ic <= Ly.imx - 1 && jc <= Ly.jmx - 1)
        cell_work4<NV, VISC>(P, a, smem + S::OFF_X, lane, r, ic, jc, k, need_dt, k_active, Flo, Fhi, pn, pn + (T_NRM - T_PK));
      if (a.want_norms) { tm_st8(tbase + 2 * T_NRM, pn + (T_NRM - T_PK)); tm_wait_st(); }
    }
    tm_fence_before();
    PT_MARK(2)
    bar_all();   // ---- records of plane k+2 are complete; the exchange area and the packets are free again
    tm_fence_after();
    PT_MARK(3)
  }
  PT_FLUSH

  if (a.want_norms) {   // per-CTA partial: warp shuffle inside the four warps that did the cell work, then across them
    double* const sred = smem + S::OFF_RED;   // [NV+1][4]
    const int crow = cell_row_of_warp(wid);
    if (crow >= 0) {
      double x[8];
      tm_ld8(tbase + 2 * T_NRM, x);
#pragma unroll
      for (int v = 0; v <= NV; ++v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x[v] += __shfl_down_sync(0xffffffffu, x[v], o);
      }
      if (lane == 0) {
#pragma unroll
        for (int v = 0; v <= NV; ++v) sred[v * 4 + crow] = x[v];
      }
    }
    bar_all();
    if (tid <= NV) {
      const double x = (sred[tid * 4] + sred[tid * 4 + 1]) + (sred[tid * 4 + 2] + sred[tid * 4 + 3]);
      const long long cta = blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z);
      a.red[cta * (NV + 1) + tid] = x;
    }
  }
  tm_fence_before();
  __syncthreads();
  if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*reinterpret_cast<volatile unsigned*>(tm_slot)), "n"(T_COLS) : "memory");
}

// k planes per CTA: long enough to amortise the prologue (two planes of gradient work), short enough for >= ~4 waves of CTAs
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
    cudaError_t e = cudaFuncSetAttribute(k_fused<NV, INTERP, SCHEME, VISC, RARE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    if (e != cudaSuccess) return F3D_ERR_CUDA;
    attr_set[ctx->device & 63] = true;
  }
  if (!ctx->tmaps_ok) return F3D_ERR_CUDA;
  TMaps tm;
  tm.q = (a.q == ctx->tm_q_ptr[0]) ? ctx->tm_q[0] : ctx->tm_q[1];
  if (a.q != ctx->tm_q_ptr[0] && a.q != ctx->tm_q_ptr[1]) return F3D_ERR_ARGUMENT;
  tm.temp = ctx->tm_temp; tm.geo = ctx->tm_geo;
  k_fused<NV, INTERP, SCHEME, VISC, RARE><<<grid, NT, shm, ctx->stream>>>(ctx->P, a, tm);
  ctx->launches++;
  return 0;
}

template <int NV, bool VISC, bool RARE>
static int launch_interp(Ctx* ctx, KArgs& a) {
  switch (ctx->P.interpolant) {
    case F3D_INTERP_NONE: return launch_one<NV, F3D_INTERP_NONE, -1, VISC, RARE>(ctx, a);
    case F3D_MUSCL:
      if (!RARE && ctx->P.scheme == F3D_AUSM) return launch_one<NV, F3D_MUSCL, F3D_AUSM, VISC, RARE>(ctx, a);   // the headline configuration
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
static int launch_fused_set(Ctx* ctx, KArgs& a) {
  if (ctx->P.sa) return ctx->P.viscous ? launch_interp<6, true, RARE>(ctx, a) : F3D_ERR_UNSUPPORTED;   // sa needs mu_ref /= 0
  if (ctx->P.viscous) return ctx->P.sst ? launch_interp<7, true, RARE>(ctx, a) : launch_interp<5, true, RARE>(ctx, a);
  return launch_interp<5, false, RARE>(ctx, a);
}

}  // namespace g4
}  // namespace f3d
