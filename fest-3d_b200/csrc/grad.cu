// Green-Gauss cell gradients of (u,v,w,T[,k,omega]) on cells 0..imx x 0..jmx x 0..kmx fused with the molecular
// (Sutherland) and eddy viscosity / SST blending function F1 of the same cell, then the ghost-gradient rule and the
// ghost mu_t / F1 copies on physical faces.
//
// Reference: src/gradients.f90:276-402 (evaluate_all_gradients), :405-482 (compute_gradient_G), :486-676
// (apply_gradient_bc, incl. the Ifaces-shaped dummy that mis-indexes Jfaces/Kfaces -- those records are gathered on the
// host into ctx->gbc with the reference's linear offset); src/viscosity.f90:109-138 (Sutherland), :343-388 (sst),
// :215-263 (sst2003), :408-465 (ghost mu_t / F1).
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
template <int NG>
__global__ void __launch_bounds__(128, 5) k_gradients(const Params P, const double* __restrict__ q, const double* __restrict__ temp,
                                                   const double* __restrict__ geom, double* __restrict__ grad, double* __restrict__ mu3, int* err,
                                                   int mode) {
  const Layout& L = P.L;
  // a warp covers cells i = 32 b - 15 .. 32 b + 16: cell 1 of a row starts a 128-byte line (ctx.hpp), so every row segment a warp
  // loads or stores is two whole lines (starting the warps at cell 0 made it three, two of them partial)
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) - G_ALIGN;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i < 0 || i > L.imx || j > L.jmx) return;
  // mode 1: only the cells whose whole stencil is interior (no ghost cell read: they can run while the halo swap and the
  // boundary fill are still in flight); mode 2: the rest; mode 0: all cells
  if (mode) {
    const bool inner = i >= 2 && i <= L.imx - 2 && j >= 2 && j <= L.jmx - 2 && k >= 2 && k <= L.kmx - 2;
    if (inner != (mode == 1)) return;
  }
  const long long fs = L.fs, c = L.idx(i, j, k), sj = L.sj, sk = L.sk;
  const double* gI = geom + (long long)G_IA * fs;
  const double* gJ = geom + (long long)G_JA * fs;
  const double* gK = geom + (long long)G_KA * fs;
  // face area vectors n*A of the six faces, per direction component
  double wlo[3][3], whi[3][3];   // [face dir][component]
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    wlo[0][d] = gI[(1 + d) * fs + c]; whi[0][d] = gI[(1 + d) * fs + c + 1];
    wlo[1][d] = gJ[(1 + d) * fs + c]; whi[1][d] = gJ[(1 + d) * fs + c + sj];
    wlo[2][d] = gK[(1 + d) * fs + c]; whi[2][d] = gK[(1 + d) * fs + c + sk];
  }
  const double AIl = gI[c], AIh = gI[c + 1], AJl = gJ[c], AJh = gJ[c + sj], AKl = gK[c], AKh = gK[c + sk];
  const double ivol2 = rcp64(2 * geom[(long long)G_VOL * fs + c]);
  const bool zgrad = L.kmx > 2;   // gradqp_z = 0 when kmx == 2 (gradients.f90:328-336)
  double g[NG][3];
  // face weights n*A once per face and direction (18 products), then 6 FMAs per gradient component: branch-free so the
  // NG*3 independent chains interleave
  double wl[3][3], wh[3][3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    wl[0][d] = wlo[0][d] * AIl; wl[1][d] = wlo[1][d] * AJl; wl[2][d] = wlo[2][d] * AKl;
    wh[0][d] = whi[0][d] * AIh; wh[1][d] = whi[1][d] * AJh; wh[2][d] = whi[2][d] * AKh;
  }
  double nan_probe = 0.0;
#pragma unroll
  for (int cc = 0; cc < NG; ++cc) {
    const double* __restrict__ var = (cc < 3) ? (q + (long long)(cc + 1) * fs) : (cc == 3 ? temp : (q + (long long)(cc + 1) * fs));
    const double v0 = var[c];
    const double sIl = var[c - 1] + v0, sJl = var[c - sj] + v0, sKl = var[c - sk] + v0;
    const double sIh = var[c + 1] + v0, sJh = var[c + sj] + v0, sKh = var[c + sk] + v0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      double r = (-sIl * wl[0][d] - sJl * wl[1][d] - sKl * wl[2][d] + sIh * wh[0][d] + sJh * wh[1][d] + sKh * wh[2][d]) * ivol2;
      if (d == 2) r = zgrad ? r : 0.0;
      nan_probe += r;
      g[cc][d] = r;
      grad[(3 * cc + d) * fs + c] = r;
    }
  }
  const bool bad = isnan(nan_probe);
  if (bad) { atomicOr(err, F3D_ERR_NAN_GRADIENT); }
  // molecular viscosity on 0..imx (elsewhere it keeps mu_ref from set-up)
  double mu = mu3[c];
  if (P.mu_variation == 1) {
    const double T = q[4 * fs + c] * rcp64(q[c] * P.R_gas);
    const double tr = T / P.T_ref;   // (T/T_ref)**1.5 = tr*sqrt(tr): <= 1 ulp from pow, 8x cheaper
    mu = P.mu_ref * (tr * sqrt(tr)) * ((P.T_ref + P.Sutherland_temp) * rcp64(T + P.Sutherland_temp));
    mu3[c] = mu;
    if (isnan(mu)) atomicOr(err, F3D_ERR_NAN_VISCOSITY);
  }
  if (NG == 5) {   // Spalart-Allmaras: mu_t = rho*tv*fv1 (viscosity.f90:149-163)
    const double tv = q[5 * fs + c], density = q[c];
    const double xi = tv * density / mu;
    const double fv1 = (pow3(xi)) / ((pow3(xi)) + (pow3(kCv1)));
    mu3[fs + c] = density * tv * fv1;
  }
  if (NG == 6) {
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
    mu3[fs + c] = NUM * rcp64(DENOM);
    const double CD = dmax(2 * density * kSigmaW2 * (g[4][0] * g[5][0] + g[4][1] * g[5][1] + g[4][2] * g[5][2]) * rcp64(tw), P.mut_floor);
    const double right = 4 * (density * kSigmaW2 * tk) * rcp64(CD * (d * d));
    const double left = dmax(var1, var2);
    const double arg1 = dmin(left, right);
    mu3[2 * fs + c] = tanh((arg1 * arg1) * (arg1 * arg1));
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// TMA-staged form of the same kernel (EXPERIMENTAL, off by default: slower as it stands, see launch_gradients).  k_gradients above
// waits on memory latency (ncu: long_scoreboard 10.9 warps per issue at
// 52 % of the DRAM peak with ideal traffic; every stencil value comes through a dependent LDG and the L1 re-fetches j/k
// neighbours from L2: 6.8 GB for 3.3 GB of DRAM reads).  Here a CTA owns a 32 x 4 column of cells and marches in k; planes
// k-1 .. k+2 of q and Temp sit in a four-deep shared-memory ring filled by the TMA engine (two cp.async.bulk.tensor.4d per plane,
// issued two planes ahead, completion on the slot's mbarrier), all stencil reads are LDS, the k neighbours are the other ring
// slots, and the only global loads left are the thread's own face metrics, requested before the mbarrier wait.
namespace gt {
constexpr int TX = 32, TY = 4, PW = TX + 4, ROWS = TY + 2, PSG = PW * ROWS, NRING = 4;
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(void* b, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
}  // namespace gt

template <int NG>
__global__ void __launch_bounds__(gt::TX * gt::TY) k_gradients_tma(const Params P, const double* __restrict__ geom, double* __restrict__ grad,
                                                                    double* __restrict__ mu3, int* err, const __grid_constant__ CUtensorMap tmq,
                                                                    const __grid_constant__ CUtensorMap tmt, int kchunk) {
  using namespace gt;
  constexpr int NV = NG + 1, NVE = (NV + 1) & ~1;      // staged q fields: an even count keeps the Temp sub-box 128-byte aligned
  constexpr int PLANE = (NVE + 2) * PSG;               // q fields, Temp, one pad field (plane size a multiple of 128 bytes)
  extern __shared__ __align__(128) double sm[];
  __shared__ __align__(8) unsigned long long mbar[NRING];
  const Layout& L = P.L;
  const int tid = threadIdx.y * TX + threadIdx.x;
  const int i0 = blockIdx.x * TX, j0 = blockIdx.y * TY;
  const int kb = blockIdx.z * kchunk, ke = min(kb + kchunk, L.kmx + 1);   // cell planes kb .. ke-1 of 0 .. kmx
  const int i = i0 + threadIdx.x, j = j0 + threadIdx.y;
  const bool valid = i <= L.imx && j <= L.jmx;
  const long long fs = L.fs, sj = L.sj, sk = L.sk;
  const int s = (threadIdx.y + 1) * PW + threadIdx.x + 1;   // slot of the cell: column 0 is i0-1, row 0 is j0-1
  if (tid == 0) {
    for (int b = 0; b < NRING; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[b])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  constexpr unsigned BYTES = 8u * PW * ROWS * (NVE + 1);
  auto issue = [&](int p) {   // plane p (cell index, -1 .. kmx+1) -> ring slot (p + 1) & 3
    if (tid != 0) return;
    const int b = (p + 1) & (NRING - 1);
    double* dst = sm + b * PLANE;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar[b])), "r"(BYTES) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(dst)),
                 "l"(&tmq), "r"(i0 + 14), "r"(j0 + 1), "r"(p + 2), "r"(0), "r"(smem_u32(&mbar[b]))
                 : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                     smem_u32(dst + NVE * PSG)),
                 "l"(&tmt), "r"(i0 + 14), "r"(j0 + 1), "r"(p + 2), "r"(0), "r"(smem_u32(&mbar[b]))
                 : "memory");
  };
  // use count of a ring slot: plane p is use number (p - (kb-1)) / NRING of its slot
  auto parity = [&](int p) { return (unsigned)(((p - (kb - 1)) / NRING) & 1); };
  issue(kb - 1); issue(kb); issue(kb + 1);
  const double* gI = geom + (long long)G_IA * fs;
  const double* gJ = geom + (long long)G_JA * fs;
  const double* gK = geom + (long long)G_KA * fs;
  const bool zgrad = L.kmx > 2;   // gradqp_z = 0 when kmx == 2 (gradients.f90:328-336)
  for (int k = kb; k < ke; ++k) {
    __syncthreads();                       // everyone is done with plane k-2: its slot takes plane k+2
    if (k + 2 <= ke) issue(k + 2);
    const long long c = L.idx(valid ? i : L.imx, valid ? j : L.jmx, k);
    // own face metrics (n*A of the six faces), requested before the wait
    double wl[3][3], wh[3][3];
    {
      const double AIl = gI[c], AIh = gI[c + 1], AJl = gJ[c], AJh = gJ[c + sj], AKl = gK[c], AKh = gK[c + sk];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        wl[0][d] = gI[(1 + d) * fs + c] * AIl; wh[0][d] = gI[(1 + d) * fs + c + 1] * AIh;
        wl[1][d] = gJ[(1 + d) * fs + c] * AJl; wh[1][d] = gJ[(1 + d) * fs + c + sj] * AJh;
        wl[2][d] = gK[(1 + d) * fs + c] * AKl; wh[2][d] = gK[(1 + d) * fs + c + sk] * AKh;
      }
    }
    const double ivol2 = rcp64(2 * geom[(long long)G_VOL * fs + c]);
    mbar_wait(&mbar[k & (NRING - 1)], parity(k - 1));           // plane k-1
    mbar_wait(&mbar[(k + 1) & (NRING - 1)], parity(k));         // plane k
    mbar_wait(&mbar[(k + 2) & (NRING - 1)], parity(k + 1));     // plane k+1
    if (!valid) continue;
    const double* pm = sm + (k & (NRING - 1)) * PLANE + s;           // plane k-1
    const double* p0 = sm + ((k + 1) & (NRING - 1)) * PLANE + s;     // plane k
    const double* pp = sm + ((k + 2) & (NRING - 1)) * PLANE + s;     // plane k+1
    double g[NG][3];
    double nan_probe = 0.0;
#pragma unroll
    for (int cc = 0; cc < NG; ++cc) {
      const int f = (cc == 3) ? NVE : cc + 1;   // u, v, w, Temp, then the turbulence variables
      const double v0 = p0[f * PSG];
      const double sIl = p0[f * PSG - 1] + v0, sJl = p0[f * PSG - PW] + v0, sKl = pm[f * PSG] + v0;
      const double sIh = p0[f * PSG + 1] + v0, sJh = p0[f * PSG + PW] + v0, sKh = pp[f * PSG] + v0;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        double r = (-sIl * wl[0][d] - sJl * wl[1][d] - sKl * wl[2][d] + sIh * wh[0][d] + sJh * wh[1][d] + sKh * wh[2][d]) * ivol2;
        if (d == 2) r = zgrad ? r : 0.0;
        nan_probe += r;
        g[cc][d] = r;
        grad[(3 * cc + d) * fs + c] = r;
      }
    }
    if (isnan(nan_probe)) atomicOr(err, F3D_ERR_NAN_GRADIENT);
    const double density = p0[0], pres = p0[4 * PSG];
    // molecular viscosity on 0..imx (elsewhere it keeps mu_ref from set-up)
    double mu = mu3[c];
    if (P.mu_variation == 1) {
      const double T = pres * rcp64(density * P.R_gas);
      const double tr = T / P.T_ref;   // (T/T_ref)**1.5 = tr*sqrt(tr): <= 1 ulp from pow, 8x cheaper
      mu = P.mu_ref * (tr * sqrt(tr)) * ((P.T_ref + P.Sutherland_temp) * rcp64(T + P.Sutherland_temp));
      mu3[c] = mu;
      if (isnan(mu)) atomicOr(err, F3D_ERR_NAN_VISCOSITY);
    }
    if (NG == 5) {   // Spalart-Allmaras: mu_t = rho*tv*fv1 (viscosity.f90:149-163)
      const double tv = p0[5 * PSG];
      const double xi = tv * density / mu;
      const double fv1 = (pow3(xi)) / ((pow3(xi)) + (pow3(kCv1)));
      mu3[fs + c] = density * tv * fv1;
    }
    if (NG == 6) {
      const double tk = p0[5 * PSG], tw = p0[6 * PSG];
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
      mu3[fs + c] = NUM * rcp64(DENOM);
      const double CD = dmax(2 * density * kSigmaW2 * (g[4][0] * g[5][0] + g[4][1] * g[5][1] + g[4][2] * g[5][2]) * rcp64(tw), P.mut_floor);
      const double right = 4 * (density * kSigmaW2 * tk) * rcp64(CD * (d * d));
      const double left = dmax(var1, var2);
      const double arg1 = dmin(left, right);
      mu3[2 * fs + c] = tanh((arg1 * arg1) * (arg1 * arg1));
    }
  }
}

template <int NG>
static int launch_gradients_tma(Ctx* ctx) {
  using namespace gt;
  const Layout& L = ctx->P.L;
  constexpr int NV = NG + 1, NVE = (NV + 1) & ~1;
  const size_t shm = sizeof(double) * NRING * (NVE + 2) * PSG;
  static bool attr_set[64] = {false};
  if (!attr_set[ctx->device & 63]) {
    if (cudaFuncSetAttribute(k_gradients_tma<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm) != cudaSuccess) return F3D_ERR_CUDA;
    attr_set[ctx->device & 63] = true;
  }
  const int nk = L.kmx + 1;
  const long long tiles = (long long)((L.imx + 1 + TX - 1) / TX) * ((L.jmx + 1 + TY - 1) / TY);
  int chunk = nk;
  while (chunk > 16 && tiles * ((nk + chunk - 1) / chunk) < 148 * 3 * 6) chunk = (chunk + 1) / 2;   // >= ~6 waves of 3 CTAs per SM
  dim3 block(TX, TY, 1), grid((L.imx + 1 + TX - 1) / TX, (L.jmx + 1 + TY - 1) / TY, (nk + chunk - 1) / chunk);
  const CUtensorMap& tq = (ctx->qp == ctx->tm_q_ptr[0]) ? ctx->tm_qg[0] : ctx->tm_qg[1];
  k_gradients_tma<NG><<<grid, block, shm, ctx->stream>>>(ctx->P, ctx->geom, ctx->grad, ctx->mu, ctx->err_dev, tq, ctx->tm_temp, chunk);
  return 0;
}

// ghost-gradient rule + ghost mu_t/F1 on one physical face (gradients.f90:638-674, viscosity.f90:408-465)
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
  double qI[NG], qG[NG], gI[NG][3];
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
    grad[(3 * cc + 0) * fs + cg] = gx + (gIx - dot * nx);
    grad[(3 * cc + 1) * fs + cg] = gy + (gIy - dot * ny);
    grad[(3 * cc + 2) * fs + cg] = gz + (gIz - dot * nz);
  }
  if (NG == 5) {   // viscosity.f90:165-212
    if (id == -5) mu3[fs + cg] = -mu3[fs + ci];
    else if (id == -1 || id == -2 || id == -3 || id == -4 || id == -6 || id == -7 || id == -8 || id == -9) mu3[fs + cg] = mu3[fs + ci];
  }
  if (NG == 6) {
    if (id == -5) { mu3[fs + cg] = -mu3[fs + ci]; mu3[2 * fs + cg] = mu3[2 * fs + ci]; }
    else if (id == -1 || id == -2 || id == -3 || id == -4 || id == -6 || id == -7 || id == -8 || id == -9) {
      mu3[fs + cg] = mu3[fs + ci]; mu3[2 * fs + cg] = mu3[2 * fs + ci];
    }
  }
}

int launch_gradients(Ctx* ctx, int mode) {
  const Layout& L = ctx->P.L;
  dim3 block(32, 4, 1);
  dim3 grid((L.imx + 1 + G_ALIGN + 31) / 32, (L.jmx + 1 + 3) / 4, L.kmx + 1);
  static int use_tma = -1;   // F3D_GRAD_TMA=1 selects the TMA-staged kernel: measured 1.88 ms against 1.47 ms for the one-thread-per-cell
                             // kernel at 256^3 (its 69 KB ring leaves 12 warps per SM and the face metrics are still plain loads)
  if (use_tma < 0) { const char* e = getenv("F3D_GRAD_TMA"); use_tma = (e && e[0] == '1') ? 1 : 0; }
  if (use_tma && mode == 0 && ctx->tmaps_ok) {
    int rc = ctx->P.sa ? launch_gradients_tma<5>(ctx) : (ctx->P.sst ? launch_gradients_tma<6>(ctx) : launch_gradients_tma<4>(ctx));
    if (rc) return rc;
  } else if (ctx->P.sa) k_gradients<5><<<grid, block, 0, ctx->stream>>>(ctx->P, ctx->qp, ctx->temp, ctx->geom, ctx->grad, ctx->mu, ctx->err_dev, mode);
  else if (ctx->P.sst) k_gradients<6><<<grid, block, 0, ctx->stream>>>(ctx->P, ctx->qp, ctx->temp, ctx->geom, ctx->grad, ctx->mu, ctx->err_dev, mode);
  else k_gradients<4><<<grid, block, 0, ctx->stream>>>(ctx->P, ctx->qp, ctx->temp, ctx->geom, ctx->grad, ctx->mu, ctx->err_dev, mode);
  ctx->launches++;
  if (mode == 1) { F3D_CUDA(cudaGetLastError()); return 0; }   // the ghost rules follow the second part
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
    if (ctx->P.sa) k_gradient_bc<5><<<g2, block, 0, ctx->stream>>>(ctx->P, ctx->qp, ctx->temp, ctx->geom, ctx->grad, ctx->mu, ctx->gbc, ctx->gbc_off_dev, mask);
    else if (ctx->P.sst) k_gradient_bc<6><<<g2, block, 0, ctx->stream>>>(ctx->P, ctx->qp, ctx->temp, ctx->geom, ctx->grad, ctx->mu, ctx->gbc, ctx->gbc_off_dev, mask);
    else k_gradient_bc<4><<<g2, block, 0, ctx->stream>>>(ctx->P, ctx->qp, ctx->temp, ctx->geom, ctx->grad, ctx->mu, ctx->gbc, ctx->gbc_off_dev, mask);
    ctx->launches++;
  }
  F3D_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace f3d
