// Implicit LU-SGS update (time_step_accuracy = implicit) for laminar / inviscid flow, the SST models, Spalart-Allmaras and k-kL: SURVEY.md
// 8(f) rank 4 (laminar) and the other four routines of the reference's dispatcher (SST, k-kL, SA, SST + lctm2015).
//
// Reference: src/lusgs.f90:134-183 (update_with_lusgs), :186-488 (update_laminar_variables), :491-630 (Flux), :633-683
// (SpectralRadius), :686-1024 (update_SST_variables), :1027-1196 (SSTFlux), :1198-1512 (update_KKL_variables), :1515-1677 (KKLFlux),
// :1680-2101 (update_SA_variables), :2104-2259 (SAFlux); called from update.f90:216-219 after one residual evaluation and the time-step
// computation.
//
// The reference sweeps the block lexicographically (forward k,j,i ascending; backward descending): cell (i,j,k) reads the correction of
// (i-1,j,k), (i,j-1,k), (i,j,k-1) in the forward sweep and of the three high neighbours in the backward one.  Cells on a hyperplane
// i+j+k = h do not depend on each other, and every cell reads only finished values whichever order the hyperplanes' cells are taken in,
// so marching the hyperplanes h = 3 .. imx+jmx+kmx-3 (one launch each, ascending then descending) reproduces the reference bit for bit up
// to FMA contraction -- there is no reduction whose order could change.  Corrections outside the block are zero (delQ / delQstar are
// allocated 0:imx and never written there, lusgs.f90:114-115): the blocks of a multi-block case are not coupled inside the sweeps.
//
// Work shared by both sweeps is done once, in a data-parallel pre-pass: the spectral radius times area of every face (SpectralRadius is
// symmetric in its two cells bit for bit: sums commute, |x| = |-x|), three face fields.
//
// Kernels: k_lusgs_lambda (pre-pass), k_lusgs_sweep<NV, FWD> (one hyperplane, four lanes per cell), k_lusgs_apply<NV> (conservative update,
// in place).  Access along a hyperplane is strided (every cell of a warp sits in its own row): this path is latency- / sector-bound, not
// the benchmark path; DESIGN.md section 3.6 has the measured cost.
#include "ctx.hpp"
#include "physics.cuh"
#include <algorithm>

namespace f3d {

namespace {

struct LFace { double A, nx, ny, nz, vol, mmu, tmu, F1; };
enum { M_LAM = 0, M_SST = 1, M_SA = 2, M_KKL = 3, M_LCTM = 4 };   // which routine of lusgs.f90: :186 laminar, :686 SST, :1680 SA, :1198 k-kL, :2262 lctm2015

// lusgs.f90:491-630 (Flux) / :1027-1196 (SSTFlux) / :1515-1677 (KKLFlux) / :2104-2259 (SAFlux)
template <int NV, int MODEL>
__device__ __forceinline__ void lusgs_flux(const Params& P, const double (&ql)[NV], const double (&qr)[NV], const double (&du)[NV], const LFace& f,
                                           double (&Flux)[NV]) {
  const double gm = P.gm, R_gas = P.R_gas;
  double U[NV], W[NV];
  U[0] = ql[0];
  U[1] = ql[0] * ql[1];
  U[2] = ql[0] * ql[2];
  U[3] = ql[0] * ql[3];
  U[4] = (ql[4] / (gm - 1.0)) + (0.5 * ql[0] * (((ql[1] * ql[1]) + (ql[2] * ql[2])) + (ql[3] * ql[3])));
  if (NV >= 6) U[5] = ql[0] * ql[5];
  if (NV >= 7) U[6] = ql[0] * ql[6];
  if (NV == 8) U[7] = ql[0] * ql[7];
#pragma unroll
  for (int l = 0; l < NV; ++l) U[l] = U[l] + du[l];
  // Divisions by one and the same denominator are taken as multiplications with its reciprocal (U(1), Volume: 6 + 21 IEEE divisions per
  // flux in the reference's text, 12 fluxes per cell and iteration; each is a ~40-instruction sequence on this machine and the sweeps are
  // bound by the length of one thread's instruction stream).  <= 1 ulp per quotient, inside the 1e-10 history tolerance like rcp64 in the sweep.
  const double iU0 = 1.0 / U[0];
  W[0] = U[0];
  W[1] = U[1] * iU0;
  W[2] = U[2] * iU0;
  W[3] = U[3] * iU0;
  W[4] = (gm - 1.0) * (U[4] - (0.5 * (((U[1] * U[1]) + (U[2] * U[2])) + (U[3] * U[3])) * iU0));
  if (MODEL == M_SST || MODEL == M_LCTM) {
    W[5] = U[5] * iU0;
    W[6] = U[6] * iU0;
    W[5] = W[5] + 0.5 * (1. - copysign(1.0, W[5])) * (ql[5] - W[5]);
    W[6] = W[6] + 0.5 * (1. - copysign(1.0, W[6])) * (ql[6] - W[6]);
    if (MODEL == M_LCTM) W[NV - 1] = fmax(U[NV - 1] * iU0, 0.0);   // lusgs.f90:2715
  }
  if (MODEL == M_KKL) { W[5] = fmax(U[5] * iU0, 1e-8); W[6] = fmax(U[6] * iU0, 1e-8); }   // lusgs.f90:1551-1554
  if (MODEL == M_SA) W[5] = fmax(U[5] * iU0, 1e-8);                                          // lusgs.f90:2140-2141
  const double nx = f.nx, ny = f.ny, nz = f.nz, Area = f.A, mmu = f.mmu, tmu = f.tmu;
  const double FaceNormalVelocity = (W[1] * nx) + (W[2] * ny) + (W[3] * nz);
  const double uface = 0.5 * (W[1] + qr[1]), vface = 0.5 * (W[2] + qr[2]), wface = 0.5 * (W[3] + qr[3]);
  Flux[0] = W[0] * FaceNormalVelocity;
  Flux[1] = (W[1] * Flux[0]) + (W[4] * nx);
  Flux[2] = (W[2] * Flux[0]) + (W[4] * ny);
  Flux[3] = (W[3] * Flux[0]) + (W[4] * nz);
  const double HalfRhoUsquare = 0.5 * W[0] * (W[1] * W[1] + W[2] * W[2] + W[3] * W[3]);
  const double RhoHt = ((gm / (gm - 1.0)) * W[4]) + HalfRhoUsquare;
  Flux[4] = RhoHt * FaceNormalVelocity;
  if (NV >= 6) Flux[5] = (W[5] * Flux[0]);
  if (NV >= 7) Flux[6] = (W[6] * Flux[0]);
  if (NV == 8) Flux[NV - 1] = (W[NV - 1] * Flux[0]);
  const double muCap = (MODEL == M_SA) ? 0.25 * (qr[0] + W[0]) * (qr[5] + W[5]) : 0.0;   // lusgs.f90:2155
  const double mu = mmu + tmu;
  const double T1 = W[4] / (W[0] * R_gas), T2 = qr[4] / (qr[0] * R_gas);
  const double iV = 1.0 / f.vol;
  const double ax = nx * Area * iV, ay = ny * Area * iV, az = nz * Area * iV;   // n A / V of the one-sided differences
  const double dTdx = (T2 - T1) * ax, dTdy = (T2 - T1) * ay, dTdz = (T2 - T1) * az;
  const double dudx = (qr[1] - W[1]) * ax, dudy = (qr[1] - W[1]) * ay, dudz = (qr[1] - W[1]) * az;
  const double dvdx = (qr[2] - W[2]) * ax, dvdy = (qr[2] - W[2]) * ay, dvdz = (qr[2] - W[2]) * az;
  const double dwdx = (qr[3] - W[3]) * ax, dwdy = (qr[3] - W[3]) * ay, dwdz = (qr[3] - W[3]) * az;
  const double trace = dudx + dvdy + dwdz;
  const double tr3 = trace * (1.0 / 3.0);
  const double Tauxx = 2. * mu * (dudx - tr3), Tauyy = 2. * mu * (dvdy - tr3), Tauzz = 2. * mu * (dwdz - tr3);
  const double Tauxy = mu * (dvdx + dudy), Tauxz = mu * (dwdx + dudz), Tauyz = mu * (dwdy + dvdz);
  const double K_heat = (mmu * P.inv_Pr + tmu * P.inv_tPr) * gm * R_gas * P.inv_gm1;
  const double Qx = K_heat * dTdx, Qy = K_heat * dTdy, Qz = K_heat * dTdz;
  Flux[1] = Flux[1] - (Tauxx * nx + Tauxy * ny + Tauxz * nz);
  Flux[2] = Flux[2] - (Tauxy * nx + Tauyy * ny + Tauyz * nz);
  Flux[3] = Flux[3] - (Tauxz * nx + Tauyz * ny + Tauzz * nz);
  Flux[4] = Flux[4] - (Tauxx * uface + Tauxy * vface + Tauxz * wface + Qx) * nx;
  Flux[4] = Flux[4] - (Tauxy * uface + Tauyy * vface + Tauyz * wface + Qy) * ny;
  Flux[4] = Flux[4] - (Tauxz * uface + Tauyz * vface + Tauzz * wface + Qz) * nz;
  if (MODEL == M_SA) {   // lusgs.f90:2171-2173, 2192
    const double dtvdx = (qr[5] - W[5]) * ax, dtvdy = (qr[5] - W[5]) * ay, dtvdz = (qr[5] - W[5]) * az;
    Flux[5] = Flux[5] + (mmu + muCap) * (dtvdx * nx + dtvdy * ny + dtvdz * nz) / kSigmaSA;
  }
  if (NV >= 7) {
    const double dtkdx = (qr[5] - W[5]) * ax, dtkdy = (qr[5] - W[5]) * ay, dtkdz = (qr[5] - W[5]) * az;
    const double dtwdx = (qr[6] - W[6]) * ax, dtwdy = (qr[6] - W[6]) * ay, dtwdz = (qr[6] - W[6]) * az;
    const double sigma_k = (MODEL == M_KKL) ? 1.0 : kSigmaK1 * f.F1 + kSigmaK2 * (1.0 - f.F1);   // global_kkl.f90: sigma_k = sigma_phi = 1
    const double sigma_w = (MODEL == M_KKL) ? 1.0 : kSigmaW1 * f.F1 + kSigmaW2 * (1.0 - f.F1);
    Flux[5] = Flux[5] + (mmu + sigma_k * tmu) * (dtkdx * nx + dtkdy * ny + dtkdz * nz);
    Flux[6] = Flux[6] + (mmu + sigma_w * tmu) * (dtwdx * nx + dtwdy * ny + dtwdz * nz);
  }
  if (NV == 8) {   // lusgs.f90:2753-2755, 2777
    const double dg = qr[NV - 1] - W[NV - 1];
    Flux[NV - 1] = Flux[NV - 1] + (mmu + tmu) * ((dg * ax) * nx + (dg * ay) * ny + (dg * az) * nz);
  }
#pragma unroll
  for (int l = 0; l < NV; ++l) Flux[l] = Flux[l] * Area;
}

// SpectralRadius (lusgs.f90:633-683) of the low face of cell (i,j,k) in each direction, into lam[d] at the cell's index (= the face's)
template <int NV, int MODEL>
__global__ void __launch_bounds__(128) k_lusgs_lambda(const Params P, const double* __restrict__ q, const double* __restrict__ geom,
                                                      const double* __restrict__ mu3 /* mu [, mu_t, F1] or nullptr */, double* __restrict__ lam) {
  const Layout& L = P.L;
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y * blockDim.y + threadIdx.y, k = 1 + blockIdx.z;
  if (i > L.imx || j > L.jmx) return;
  const long long fs = L.fs, c = L.idx(i, j, k);
  const long long st[3] = {1, L.sj, L.sk};
  const int pos[3] = {i, j, k}, mx[3] = {L.imx, L.jmx, L.kmx};
  const double cx = geom[(long long)G_CX * fs + c], cy = geom[(long long)G_CY * fs + c], cz = geom[(long long)G_CZ * fs + c];
  const double r0 = q[c], u0 = q[fs + c], v0 = q[2 * fs + c], w0 = q[3 * fs + c], p0 = q[4 * fs + c];
  const double m0 = mu3 ? mu3[c] : 0.0, t0 = (mu3 && MODEL != M_LAM) ? mu3[fs + c] : 0.0;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    // the face exists when the cell index is interior in the two other directions
    bool ok = true;
#pragma unroll
    for (int e = 0; e < 3; ++e) if (e != d && pos[e] > mx[e] - 1) ok = false;
    if (!ok) continue;
    const long long n = c - st[d];
    const double* gf = geom + (long long)(G_IA + 4 * d) * fs;
    const double A = gf[c], nx = gf[fs + c], ny = gf[2 * fs + c], nz = gf[3 * fs + c];
    const double r1 = q[n], u1 = q[fs + n], v1 = q[2 * fs + n], w1 = q[3 * fs + n], p1 = q[4 * fs + n];
    double NormalSpeed = 0.5 * (((u1 + u0) * nx) + ((v1 + v0) * ny) + ((w1 + w0) * nz));
    NormalSpeed = fabs(NormalSpeed);
    const double SpeedOfSound = 0.5 * (sqrt(P.gm * p1 / r1) + sqrt(P.gm * p0 / r0));
    const double rho = 0.5 * (r1 + r0);
    const double dx = geom[(long long)G_CX * fs + n] - cx, dy = geom[(long long)G_CY * fs + n] - cy, dz = geom[(long long)G_CZ * fs + n] - cz;
    const double distance = sqrt(((dx * dx) + (dy * dy)) + (dz * dz));
    const double mm = mu3 ? 0.5 * (mu3[n] + m0) : 0.0, tm = (mu3 && MODEL != M_LAM) ? 0.5 * (mu3[fs + n] + t0) : 0.0;
    const double vis = P.gm * (mm / P.Pr + tm / P.tPr) / (rho * distance);
    lam[d * fs + c] = (NormalSpeed + SpeedOfSound + vis) * A;
  }
}

// One hyperplane i + j + k = h of a sweep.  FWD: delQstar from the low neighbours (lusgs.f90:296-312, 802-838); else delQ from the high
// neighbours (:426-447, 920-959).  Grid: x over i (32 cells per CTA), y over the k planes the hyperplane crosses (k = k0 + blockIdx.y).
// FOUR lanes per cell: lanes 0..2 evaluate the flux pair of the I, J, K face, lane 3 the diagonal D; the three face terms meet in lane 0
// by shuffles and are added in the reference's order ((I) + (J)) + (K).  A sweep is bound by the length of one thread's instruction stream
// times the number of hyperplanes (a hyperplane of a 128^3 block holds <= 12 k cells, one of SmoothBump 48): splitting the cell's work
// across lanes shortens that stream.  Measured against one thread per cell: the reference's three shipped cases 80 -> 15 s, 128^3 9.0 -> 8.6 ms,
// 256^3 SST 63.5 -> 54.4 ms per iteration (profiles/r02_lusgs_timing.txt).
template <int NV, int MODEL, bool FWD>
__global__ void __launch_bounds__(128) k_lusgs_sweep(const Params P, const double* __restrict__ q, const double* __restrict__ geom,
                                                      const double* __restrict__ mu3, const double* __restrict__ lam, const double* __restrict__ dt,
                                                      const double* __restrict__ residue, double* __restrict__ dqs, double* __restrict__ dq,
                                                      const double* __restrict__ grad /* sa: vorticity of the source Jacobian */, int h, int k0) {
  const Layout& L = P.L;
  const int k = k0 + blockIdx.y;
  const int lane4 = threadIdx.x & 3;
  const int i = 1 + blockIdx.x * (blockDim.x >> 2) + (threadIdx.x >> 2);
  const int j = h - i - k;
  const bool valid = !(i > L.imx - 1 || j < 1 || j > L.jmx - 1);   // no early return: every lane of the warp takes part in the shuffles
  const long long fs = L.fs, c = valid ? L.idx(i, j, k) : L.idx(1, 1, 1);
  const long long st[3] = {1, L.sj, L.sk};
  double term[NV], D[NV];
#pragma unroll
  for (int l = 0; l < NV; ++l) { term[l] = 0.0; D[l] = 1.0; }
  if (valid && lane4 < 3) {
    const int d = lane4;
    double Q0[NV];
#pragma unroll
    for (int l = 0; l < NV; ++l) Q0[l] = q[l * fs + c];
    const double vol0 = geom[(long long)G_VOL * fs + c];
    const double m0 = mu3 ? mu3[c] : 0.0, t0 = (mu3 && MODEL != M_LAM) ? mu3[fs + c] : 0.0,
                 f0 = (mu3 && (MODEL == M_SST || MODEL == M_LCTM)) ? mu3[2 * fs + c] : 0.0;
    const long long n = FWD ? c - st[d] : c + st[d];
    const long long fc = FWD ? c : c + st[d];          // index of the face record
    const double sg = FWD ? -1.0 : 1.0;
    const double* gf = geom + (long long)(G_IA + 4 * d) * fs;
    LFace f;
    f.A = gf[fc]; f.nx = sg * gf[fs + fc]; f.ny = sg * gf[2 * fs + fc]; f.nz = sg * gf[3 * fs + fc];
    f.vol = 0.5 * (geom[(long long)G_VOL * fs + n] + vol0);
    f.mmu = mu3 ? 0.5 * (mu3[n] + m0) : 0.0;
    f.tmu = (mu3 && MODEL != M_LAM) ? 0.5 * (mu3[fs + n] + t0) : 0.0;
    f.F1 = (mu3 && (MODEL == M_SST || MODEL == M_LCTM)) ? 0.5 * (mu3[2 * fs + n] + f0) : 0.0;
    const double* __restrict__ src = FWD ? dqs : dq;
    double Qn[NV], DQ[NV], zero[NV], Fn[NV], Fo[NV];
#pragma unroll
    for (int l = 0; l < NV; ++l) { Qn[l] = q[l * fs + n]; DQ[l] = src[l * fs + n]; zero[l] = 0.0; }
    lusgs_flux<NV, MODEL>(P, Qn, Q0, DQ, f, Fn);
    lusgs_flux<NV, MODEL>(P, Qn, Q0, zero, f, Fo);
    const double lm = lam[d * fs + fc];
#pragma unroll
    for (int l = 0; l < NV; ++l) term[l] = ((Fn[l] - Fo[l]) - lm * DQ[l]);
  } else if (valid) {   // lane 3: D = V / dt + sum(lambda A) / 2 (+ the SST source Jacobian)
    double s = 0.0;   // LambdaTimesArea(1..6): low I, J, K faces, then high I, J, K faces; SUM in that order
#pragma unroll
    for (int d = 0; d < 3; ++d) s = s + lam[d * fs + c];
#pragma unroll
    for (int d = 0; d < 3; ++d) s = s + lam[d * fs + c + st[d]];
    const double vol0 = geom[(long long)G_VOL * fs + c];
    const double D0 = (vol0 / dt[c]) + 0.5 * s;
#pragma unroll
    for (int l = 0; l < NV; ++l) D[l] = D0;
    if (MODEL == M_SST || MODEL == M_LCTM) {   // lusgs.f90:830-832, 2406-2409
      const double f0 = mu3 ? mu3[2 * fs + c] : 0.0, tw = q[6 * fs + c];
      const double beta = f0 * kBeta1 + (1.0 - f0) * kBeta2;
      D[5] = (D[5] + (kBstar * tw) * vol0);
      D[6] = (D[6] + 2.0 * beta * tw * vol0);
    }
    if (MODEL == M_LCTM) {   // lusgs.f90:2410-2440: derivative of the intermittency source (no pressure-gradient factor in this Re_theta)
      const double density = q[c], tk = q[5 * fs + c], tw = q[6 * fs + c], gm_ = q[7 * fs + c], d = geom[(long long)G_DIST * fs + c], muc = mu3[c];
      double g[3][3];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) { g[cc][0] = grad[(3 * cc) * fs + c]; g[cc][1] = grad[(3 * cc + 1) * fs + c]; g[cc][2] = grad[(3 * cc + 2) * fs + c]; }
      const double wx = g[2][1] - g[1][2], wy = g[0][2] - g[2][0], wz = g[1][0] - g[0][1];
      const double vort = sqrt(wx * wx + wy * wy + wz * wz);
      const double syz = g[2][1] + g[1][2], szx = g[0][2] + g[2][0], sxy = g[1][0] + g[0][1];
      const double strain = sqrt(((syz * syz) + (szx * szx) + (sxy * sxy) + 2 * (g[0][0] * g[0][0]) + 2 * (g[1][1] * g[1][1]) + 2 * (g[2][2] * g[2][2])));
      const double TuL = fmin(100.0 * sqrt(2.0 * tk / 3.0) / (tw * d), 100.0);
      const double Re_theta = 100.0 + 1000.0 * exp(-TuL);
      const double Rev = density * d * d * strain / muc;
      const double RT = density * tk / (muc * tw);
      const double hr = 0.5 * RT;
      const double Fturb = exp(-((hr * hr) * (hr * hr)));
      const double Fonset2 = fmin(Rev / (2.2 * Re_theta), 2.0);
      const double r35 = RT / 3.5;
      const double Fonset3 = fmax(1.0 - (r35 * r35 * r35), 0.0);
      const double Fonset = fmax(Fonset2 - Fonset3, 0.0);
      const double Dp = 100 * density * strain * Fonset * (1.0 - 2.0 * gm_);
      const double De = 0.06 * vort * Fturb * density * (2.0 * 50.0 * gm_ - 1.0);
      D[NV - 1] = (D[NV - 1] + (-Dp + De) * vol0);
    }
    if (MODEL == M_KKL) {   // lusgs.f90:1339-1341
      const double rho = q[c], tk = q[5 * fs + c], tkl = q[6 * fs + c], d = geom[(long long)G_DIST * fs + c], mu_c = mu3[c];
      D[5] = D[5] + (2.5 * kKklCmu75 * rho * (tk * sqrt(tk)) * vol0 / tkl);
      D[5] = D[5] + (2 * mu_c * vol0 / (d * d));
      D[6] = D[6] + (6 * mu_c * vol0 / (d * d));
    }
    if (MODEL == M_SA) {   // lusgs.f90:1868-1913: the source-term derivatives go into EVERY component of D (array assignment) -- reproduced
      const double density = q[c], tv = q[5 * fs + c];
      const double a = grad[(3 * 2 + 1) * fs + c] - grad[(3 * 1 + 2) * fs + c], b = grad[(3 * 0 + 2) * fs + c] - grad[(3 * 2 + 0) * fs + c],
                   cc = grad[(3 * 1 + 0) * fs + c] - grad[(3 * 0 + 1) * fs + c];
      const double Omega = sqrt(((a * a) + (b * b) + (cc * cc)));
      const double dist_i = geom[(long long)G_DIST * fs + c], dist_i_2 = dist_i * dist_i, k2 = kKappaSA * kKappaSA;
      const double nu = mu3[c] / density;
      const double Ji = tv / nu, Ji_2 = Ji * Ji, Ji_3 = Ji_2 * Ji;
      const double cv1_3 = pow3(kCv1), cw3_6 = pow6(kCw3);
      const double fv1 = (Ji_3) / ((Ji_3) + (cv1_3));
      const double fv2 = 1.0 - Ji / (1.0 + (Ji * fv1));
      const double inv_k2_d2 = 1.0 / (k2 * dist_i_2);
      double Shat = Omega + tv * fv2 * inv_k2_d2;
      Shat = fmax(Shat, 1.0e-10);
      const double inv_Shat = 1.0 / Shat;
      const double den1 = (Ji_3 + cv1_3);
      const double dfv1 = 3.0 * Ji_2 * cv1_3 / (nu * (den1 * den1));
      const double den2 = (1.0 + Ji * fv1);
      const double dfv2 = -((1.0 / nu) - Ji_2 * dfv1) / (den2 * den2);
      const double dShat = (fv2 + tv * dfv2) * inv_k2_d2;
      const double r = fmin(tv * inv_Shat * inv_k2_d2, 10.0);
      const double r2 = r * r, r6 = r2 * r2 * r2;
      const double g = r + kCw2 * ((r6) - r);
      const double g_6 = pow6(g);
      const double glim = cbrt(sqrt((1.0 + cw3_6) / (g_6 + cw3_6)));
      const double fw = g * glim;
      const double dr = (Shat - tv * dShat) * inv_Shat * inv_Shat * inv_k2_d2;
      const double dg = dr * (1.0 + kCw2 * (6.0 * (r2 * r2 * r) - 1.0));
      const double dfw = dg * glim * (1.0 - g_6 / (g_6 + cw3_6));
#pragma unroll
      for (int l = 0; l < NV; ++l) {
        D[l] = D[l] - kCb1 * (tv * dShat + Shat) * vol0;
        D[l] = D[l] + kCw1 * (dfw * tv + 2 * fw) * tv / dist_i_2 * vol0;
      }
    }
  }
#pragma unroll
  for (int l = 0; l < NV; ++l) {
    const double tJ = __shfl_down_sync(0xffffffffu, term[l], 1, 4), tK = __shfl_down_sync(0xffffffffu, term[l], 2, 4);
    const double Dl = __shfl_down_sync(0xffffffffu, D[l], 3, 4);
    if (valid && lane4 == 0) {
      const double acc = (term[l] + tJ) + tK;   // ((I) + (J)) + (K)
      if (FWD) dqs[l * fs + c] = (-residue[l * fs + c] - 0.5 * acc) / Dl;
      else dq[l * fs + c] = dqs[l * fs + c] - 0.5 * acc / Dl;
    }
  }
}

// conservative update with delQ, back to primitive variables, in place (lusgs.f90:452-486, 964-1021)
template <int NV, int MODEL>
__global__ void __launch_bounds__(128) k_lusgs_apply(const Params P, double* __restrict__ q, const double* __restrict__ dq) {
  const Layout& L = P.L;
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y * blockDim.y + threadIdx.y, k = 1 + blockIdx.z;
  if (i > L.imx - 1 || j > L.jmx - 1) return;
  const long long fs = L.fs, c = L.idx(i, j, k);
  double qq[NV], cq[NV];
#pragma unroll
  for (int l = 0; l < NV; ++l) qq[l] = q[l * fs + c];
  cq[0] = qq[0];
  cq[1] = qq[0] * qq[1];
  cq[2] = qq[0] * qq[2];
  cq[3] = qq[0] * qq[3];
  cq[4] = (qq[4] / (P.gm - 1.0)) + (0.5 * qq[0] * (((qq[1] * qq[1]) + (qq[2] * qq[2])) + (qq[3] * qq[3])));
  if (NV >= 6) cq[5] = qq[0] * qq[5];
  if (NV >= 7) cq[6] = qq[0] * qq[6];
  if (NV == 8) cq[NV - 1] = qq[0] * qq[NV - 1];
#pragma unroll
  for (int l = 0; l < NV; ++l) cq[l] = cq[l] + dq[l * fs + c];
  q[c] = cq[0];
  q[fs + c] = cq[1] / cq[0];
  q[2 * fs + c] = cq[2] / cq[0];
  q[3 * fs + c] = cq[3] / cq[0];
  q[4 * fs + c] = (P.gm - 1.0) * (cq[4] - (0.5 * (((cq[1] * cq[1]) + (cq[2] * cq[2])) + (cq[3] * cq[3])) / cq[0]));
  if (MODEL == M_SST || MODEL == M_LCTM) {
    if (cq[5] > 0) q[5 * fs + c] = cq[5] / cq[0];
    if (cq[6] > 0) q[6 * fs + c] = cq[6] / cq[0];
  }
  if (MODEL == M_LCTM) q[(NV - 1) * fs + c] = fmax(cq[NV - 1] / cq[0], 0.0);   // lusgs.f90:2663-2664: here the intermittency IS advanced
  if (MODEL == M_KKL) { q[5 * fs + c] = fmax(cq[5] / cq[0], 1.e-8); q[6 * fs + c] = fmax(cq[6] / cq[0], 1.e-8); }   // lusgs.f90:1505-1508
  if (MODEL == M_SA) q[5 * fs + c] = fmax(cq[5] / cq[0], 1.e-8);                                                     // lusgs.f90:2095-2096
}

template <int NV, int MODEL>
int lusgs_run(Ctx* ctx) {
  const Layout& L = ctx->P.L;
  const int ni = L.imx - 1, nj = L.jmx - 1, nk = L.kmx - 1;
  const double* mu3 = ctx->P.viscous ? ctx->mu : nullptr;
  cudaStream_t st = ctx->stream;
  dim3 block(32, 4, 1);
  k_lusgs_lambda<NV, MODEL><<<dim3((L.imx + 31) / 32, (L.jmx + 3) / 4, L.kmx), block, 0, st>>>(ctx->P, ctx->qp, ctx->geom, mu3, ctx->lusgs_lam);
  ctx->launches++;
  const dim3 sb(128, 1, 1);
  const int gx = (ni + 31) / 32;   // four lanes per cell
  for (int pass = 0; pass < 2; ++pass) {
    for (int hh = 3; hh <= ni + nj + nk; ++hh) {
      const int h = pass == 0 ? hh : (ni + nj + nk + 3 - hh);
      const int klo = std::max(1, h - ni - nj), khi = std::min(nk, h - 2);
      if (khi < klo) continue;
      const dim3 grid(gx, khi - klo + 1, 1);
      if (pass == 0)
        k_lusgs_sweep<NV, MODEL, true><<<grid, sb, 0, st>>>(ctx->P, ctx->qp, ctx->geom, mu3, ctx->lusgs_lam, ctx->dt, ctx->residue, ctx->lusgs_dqs, ctx->lusgs_dq, ctx->grad, h, klo);
      else
        k_lusgs_sweep<NV, MODEL, false><<<grid, sb, 0, st>>>(ctx->P, ctx->qp, ctx->geom, mu3, ctx->lusgs_lam, ctx->dt, ctx->residue, ctx->lusgs_dqs, ctx->lusgs_dq, ctx->grad, h, klo);
      ctx->launches++;
    }
  }
  k_lusgs_apply<NV, MODEL><<<dim3((ni + 31) / 32, (nj + 3) / 4, nk), block, 0, st>>>(ctx->P, ctx->qp, ctx->lusgs_dq);
  ctx->launches++;
  F3D_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int launch_lusgs(Ctx* ctx) {
  if (!ctx->lusgs_dqs || !ctx->lusgs_dq || !ctx->lusgs_lam) return F3D_ERR_ARGUMENT;
  if (ctx->P.L.nv == 5) return lusgs_run<5, M_LAM>(ctx);
  if (ctx->P.L.nv == 7 && ctx->P.kkl) return lusgs_run<7, M_KKL>(ctx);
  if (ctx->P.L.nv == 7 && ctx->P.sst) return lusgs_run<7, M_SST>(ctx);
  if (ctx->P.L.nv == 6 && ctx->P.sa) return lusgs_run<6, M_SA>(ctx);
  if (ctx->P.L.nv == 8 && ctx->P.lctm) return lusgs_run<8, M_LCTM>(ctx);
  return F3D_ERR_UNSUPPORTED;
}

}  // namespace f3d
