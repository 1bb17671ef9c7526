// TEST INFRASTRUCTURE ONLY -- CPU oracle, part 3: the implicit LU-SGS update (time_step_accuracy = implicit) for laminar / inviscid
// flow, the SST models, Spalart-Allmaras and k-kL.  PARITY UNPINNED (see oracle_abi.h).
//
// Reference: src/lusgs.f90:86-131 (setup: delQ / delQstar on 0:imx, mmu / tmu aliases), :134-183 (dispatch), :186-488
// (update_laminar_variables), :491-630 (Flux), :633-683 (SpectralRadius), :686-1024 (update_SST_variables), :1027-1196 (SSTFlux),
// :1198-1512 (update_KKL_variables), :1515-1677 (KKLFlux), :1680-2101 (update_SA_variables), :2104-2259 (SAFlux).
// The sweeps are written in the reference's loop order (forward k,j,i ascending; backward i,j,k descending with k innermost); every cell
// only reads neighbours the same sweep has already passed, so any order that respects that gives the same bits.
#include "oracle_core.hpp"

namespace orc {

namespace {

struct Face {   // "Flist" of the reference: area, signed normal, mean volume, mean mu, mean mu_t [, mean F1]
  double A, nx, ny, nz, vol, mmu, tmu, F1;
};

struct LusgsCtx {
  double gm, R_gas, Pr, tPr;
};

static inline double fsign1(double b) { return std::copysign(1.0, b); }   // sign(1., b)
enum { M_LAM = 0, M_SST = 1, M_SA = 2, M_KKL = 3, M_LCTM = 4 };   // M_LCTM: the SST routine + the intermittency (n_var 8)
// global_sa.f90:6-19, global_kkl.f90:6-15
constexpr double sa_cb1 = 0.1355, sa_cb2 = 0.6220, sa_cw2 = 0.3, sa_cw3 = 2.0, sa_cv1 = 7.1, sa_sigma = 2. / 3., sa_kappa = 0.41;
static const double sa_cw1 = (sa_cb1 / (sa_kappa * sa_kappa)) + ((1 + sa_cb2) / sa_sigma);
static const double sa_cv1_3 = sa_cv1 * sa_cv1 * sa_cv1, sa_cw3_6 = (sa_cw3 * sa_cw3 * sa_cw3) * (sa_cw3 * sa_cw3 * sa_cw3);
constexpr double kkl_cmu_l = 0.09, kkl_sigma_k_l = 1.0, kkl_sigma_phi_l = 1.0;

// lusgs.f90:491-630 (n_var 5) and :1027-1196 (n_var 7): flux through the face of the neighbour state ql advanced by du, against the cell qr
template <int NV, int MODEL>
static void lusgs_flux(const LusgsCtx& X, const double* ql, const double* qr, const double* du, const Face& f, double* Flux) {
  const double gm = X.gm, R_gas = X.R_gas;
  double U[NV], W[NV];
  const double* P = qr;
  U[0] = ql[0];
  U[1] = ql[0] * ql[1];
  U[2] = ql[0] * ql[2];
  U[3] = ql[0] * ql[3];
  U[4] = (ql[4] / (gm - 1.0)) + (0.5 * ql[0] * (((ql[1] * ql[1]) + (ql[2] * ql[2])) + (ql[3] * ql[3])));
  if (NV >= 6) U[5] = ql[0] * ql[5];
  if (NV >= 7) U[6] = ql[0] * ql[6];
  if (NV == 8) U[7] = ql[0] * ql[7];
  for (int l = 0; l < NV; ++l) U[l] = U[l] + du[l];
  W[0] = U[0];
  W[1] = U[1] / U[0];
  W[2] = U[2] / U[0];
  W[3] = U[3] / U[0];
  W[4] = (gm - 1.0) * (U[4] - (0.5 * (((U[1] * U[1]) + (U[2] * U[2])) + (U[3] * U[3])) / U[0]));
  if (MODEL == M_SST || MODEL == M_LCTM) {
    W[5] = U[5] / U[0];
    W[6] = U[6] / U[0];
    if (MODEL == M_LCTM) W[7] = U[7] / U[0];
    W[5] = W[5] + 0.5 * (1. - fsign1(W[5])) * (ql[5] - W[5]);
    W[6] = W[6] + 0.5 * (1. - fsign1(W[6])) * (ql[6] - W[6]);
    if (MODEL == M_LCTM) W[7] = std::fmax(W[7], 0.0);   // lusgs.f90:2715
  }
  if (MODEL == M_KKL) {   // lusgs.f90:1551-1554
    W[5] = U[5] / U[0];
    W[6] = U[6] / U[0];
    W[5] = std::fmax(W[5], 1e-8);
    W[6] = std::fmax(W[6], 1e-8);
  }
  if (MODEL == M_SA) {    // lusgs.f90:2140-2141
    W[5] = U[5] / U[0];
    W[5] = std::fmax(W[5], 1e-8);
  }
  const double nx = f.nx, ny = f.ny, nz = f.nz, Area = f.A, Volume = f.vol, mmu = f.mmu, tmu = f.tmu;
  const double FaceNormalVelocity = (W[1] * nx) + (W[2] * ny) + (W[3] * nz);
  const double uface = 0.5 * (W[1] + P[1]);
  const double vface = 0.5 * (W[2] + P[2]);
  const double wface = 0.5 * (W[3] + P[3]);
  Flux[0] = W[0] * FaceNormalVelocity;
  Flux[1] = (W[1] * Flux[0]) + (W[4] * nx);
  Flux[2] = (W[2] * Flux[0]) + (W[4] * ny);
  Flux[3] = (W[3] * Flux[0]) + (W[4] * nz);
  const double HalfRhoUsquare = 0.5 * W[0] * (W[1] * W[1] + W[2] * W[2] + W[3] * W[3]);
  const double RhoHt = ((gm / (gm - 1.0)) * W[4]) + HalfRhoUsquare;
  Flux[4] = RhoHt * FaceNormalVelocity;
  if (NV >= 6) Flux[5] = (W[5] * Flux[0]);
  if (NV >= 7) Flux[6] = (W[6] * Flux[0]);
  if (NV == 8) Flux[7] = (W[7] * Flux[0]);
  const double muCap = (MODEL == M_SA) ? 0.25 * (P[0] + W[0]) * (P[5] + W[5]) : 0.0;   // lusgs.f90:2155
  const double mu = mmu + tmu;
  const double T1 = W[4] / (W[0] * R_gas);
  const double T2 = P[4] / (P[0] * R_gas);
  const double dTdx = (T2 - T1) * nx * Area / Volume, dTdy = (T2 - T1) * ny * Area / Volume, dTdz = (T2 - T1) * nz * Area / Volume;
  const double dudx = (P[1] - W[1]) * nx * Area / Volume, dudy = (P[1] - W[1]) * ny * Area / Volume, dudz = (P[1] - W[1]) * nz * Area / Volume;
  const double dvdx = (P[2] - W[2]) * nx * Area / Volume, dvdy = (P[2] - W[2]) * ny * Area / Volume, dvdz = (P[2] - W[2]) * nz * Area / Volume;
  const double dwdx = (P[3] - W[3]) * nx * Area / Volume, dwdy = (P[3] - W[3]) * ny * Area / Volume, dwdz = (P[3] - W[3]) * nz * Area / Volume;
  const double trace = dudx + dvdy + dwdz;
  const double Tauxx = 2. * mu * (dudx - trace / 3.0);
  const double Tauyy = 2. * mu * (dvdy - trace / 3.0);
  const double Tauzz = 2. * mu * (dwdz - trace / 3.0);
  const double Tauxy = mu * (dvdx + dudy);
  const double Tauxz = mu * (dwdx + dudz);
  const double Tauyz = mu * (dwdy + dvdz);
  const double K_heat = (mmu / X.Pr + tmu / X.tPr) * gm * R_gas / (gm - 1.0);
  const double Qx = K_heat * dTdx, Qy = K_heat * dTdy, Qz = K_heat * dTdz;
  Flux[1] = Flux[1] - (Tauxx * nx + Tauxy * ny + Tauxz * nz);
  Flux[2] = Flux[2] - (Tauxy * nx + Tauyy * ny + Tauyz * nz);
  Flux[3] = Flux[3] - (Tauxz * nx + Tauyz * ny + Tauzz * nz);
  Flux[4] = Flux[4] - (Tauxx * uface + Tauxy * vface + Tauxz * wface + Qx) * nx;
  Flux[4] = Flux[4] - (Tauxy * uface + Tauyy * vface + Tauyz * wface + Qy) * ny;
  Flux[4] = Flux[4] - (Tauxz * uface + Tauyz * vface + Tauzz * wface + Qz) * nz;
  if (MODEL == M_SA) {   // lusgs.f90:2171-2173, 2192
    const double dtvdx = (P[5] - W[5]) * nx * Area / Volume, dtvdy = (P[5] - W[5]) * ny * Area / Volume, dtvdz = (P[5] - W[5]) * nz * Area / Volume;
    Flux[5] = Flux[5] + (mmu + muCap) * (dtvdx * nx + dtvdy * ny + dtvdz * nz) / sa_sigma;
  }
  if (NV >= 7) {
    const double dtkdx = (P[5] - W[5]) * nx * Area / Volume, dtkdy = (P[5] - W[5]) * ny * Area / Volume, dtkdz = (P[5] - W[5]) * nz * Area / Volume;
    const double dtwdx = (P[6] - W[6]) * nx * Area / Volume, dtwdy = (P[6] - W[6]) * ny * Area / Volume, dtwdz = (P[6] - W[6]) * nz * Area / Volume;
    const double sigma_k = (MODEL == M_KKL) ? kkl_sigma_k_l : sigma_k1 * f.F1 + sigma_k2 * (1.0 - f.F1);
    const double sigma_w = (MODEL == M_KKL) ? kkl_sigma_phi_l : sigma_w1 * f.F1 + sigma_w2 * (1.0 - f.F1);
    Flux[5] = Flux[5] + (mmu + sigma_k * tmu) * (dtkdx * nx + dtkdy * ny + dtkdz * nz);
    Flux[6] = Flux[6] + (mmu + sigma_w * tmu) * (dtwdx * nx + dtwdy * ny + dtwdz * nz);
  }
  if (NV == 8) {   // lusgs.f90:2753-2755, 2777
    const double dgdx = (P[7] - W[7]) * nx * Area / Volume, dgdy = (P[7] - W[7]) * ny * Area / Volume, dgdz = (P[7] - W[7]) * nz * Area / Volume;
    Flux[7] = Flux[7] + (mmu + tmu) * (dgdx * nx + dgdy * ny + dgdz * nz);
  }
  for (int l = 0; l < NV; ++l) Flux[l] = Flux[l] * Area;
}

// lusgs.f90:633-683
static double spectral_radius(const LusgsCtx& X, const double* ql, const double* qr, const Face& f, const double* c1, const double* c2) {
  double NormalSpeed = 0.5 * (((ql[1] + qr[1]) * f.nx) + ((ql[2] + qr[2]) * f.ny) + ((ql[3] + qr[3]) * f.nz));
  NormalSpeed = std::fabs(NormalSpeed);
  const double SpeedOfSound = 0.5 * (std::sqrt(X.gm * ql[4] / ql[0]) + std::sqrt(X.gm * qr[4] / qr[0]));
  const double rho = 0.5 * (ql[0] + qr[0]);
  const double distance = std::sqrt(((c1[0] - c2[0]) * (c1[0] - c2[0]) + (c1[1] - c2[1]) * (c1[1] - c2[1])) + (c1[2] - c2[2]) * (c1[2] - c2[2]));
  const double vis = X.gm * (f.mmu / X.Pr + f.tmu / X.tPr) / (rho * distance);
  return (NormalSpeed + SpeedOfSound + vis) * f.A;
}

template <int NV, int MODEL>
static void lusgs_update(Block& B) {
  const int imx = B.imx, jmx = B.jmx, kmx = B.kmx;
  const OracleConfig& c = B.c;
  const LusgsCtx X{c.gm, c.R_gas, c.Pr, c.tPr};
  const bool have_mu = c.mu_ref != 0.0, have_mut = (MODEL != M_LAM);   // lusgs.f90:117-130: mmu / tmu point at a zero array otherwise
  Arr4 delQstar, delQ;
  delQstar.alloc(0, imx, 0, jmx, 0, kmx, NV);
  delQ.alloc(0, imx, 0, jmx, 0, kmx, NV);
  auto mmu = [&](int i, int j, int k) { return have_mu ? B.mu(i, j, k) : 0.0; };
  auto tmu = [&](int i, int j, int k) { return have_mut ? B.mu_t(i, j, k) : 0.0; };
  const int di[6] = {-1, 0, 0, 1, 0, 0}, dj[6] = {0, -1, 0, 0, 1, 0}, dk[6] = {0, 0, -1, 0, 0, 1};
  auto cell_setup = [&](int i, int j, int k, double (&Q)[7][NV], Face (&F)[6], double (&L)[6], double (&D)[NV]) {
    for (int l = 0; l < NV; ++l) Q[0][l] = B.qp(i, j, k, l + 1);
    double C0[3] = {B.cells.cx(i, j, k), B.cells.cy(i, j, k), B.cells.cz(i, j, k)};
    for (int n = 0; n < 6; ++n) {
      const int a = i + di[n], b = j + dj[n], cc = k + dk[n];
      for (int l = 0; l < NV; ++l) Q[n + 1][l] = B.qp(a, b, cc, l + 1);
      const Rec4& R = (n % 3 == 0) ? B.If : ((n % 3 == 1) ? B.Jf : B.Kf);
      const int fi = (n == 3) ? i + 1 : i, fj = (n == 4) ? j + 1 : j, fk = (n == 5) ? k + 1 : k;
      const double sg = (n < 3) ? -1.0 : 1.0;
      F[n].A = R.A(fi, fj, fk);
      F[n].nx = sg * R.nx(fi, fj, fk); F[n].ny = sg * R.ny(fi, fj, fk); F[n].nz = sg * R.nz(fi, fj, fk);
      F[n].vol = 0.5 * (B.cells.vol(a, b, cc) + B.cells.vol(i, j, k));
      F[n].mmu = 0.5 * (mmu(a, b, cc) + mmu(i, j, k));
      F[n].tmu = 0.5 * (tmu(a, b, cc) + tmu(i, j, k));
      F[n].F1 = (MODEL == M_SST || MODEL == M_LCTM) ? 0.5 * (B.F1(a, b, cc) + B.F1(i, j, k)) : 0.0;
      double C1[3] = {B.cells.cx(a, b, cc), B.cells.cy(a, b, cc), B.cells.cz(a, b, cc)};
      L[n] = spectral_radius(X, Q[n + 1], Q[0], F[n], C1, C0);
    }
    double s = 0.0;
    for (int n = 0; n < 6; ++n) s = s + L[n];   // SUM(LambdaTimesArea)
    const double D0 = (B.cells.vol(i, j, k) / B.delta_t(i, j, k)) + 0.5 * s;
    for (int l = 0; l < NV; ++l) D[l] = D0;
    if (MODEL == M_SST || MODEL == M_LCTM) {   // lusgs.f90:830-832, 2406-2409
      const double beta = B.F1(i, j, k) * beta1 + (1.0 - B.F1(i, j, k)) * beta2;
      D[5] = (D[5] + (bstar * B.qp(i, j, k, 7)) * B.cells.vol(i, j, k));
      D[6] = (D[6] + 2.0 * beta * B.qp(i, j, k, 7) * B.cells.vol(i, j, k));
    }
    if (MODEL == M_LCTM) {   // lusgs.f90:2410-2440: derivative of the intermittency source (no pressure-gradient factor in this Re_theta)
      const double density = B.qp(i, j, k, 1), d = B.dist(i, j, k), muc = B.mu(i, j, k);
      const double ux = B.gx(i, j, k, 1), uy = B.gy(i, j, k, 1), uz = B.gz(i, j, k, 1);
      const double vx = B.gx(i, j, k, 2), vy = B.gy(i, j, k, 2), vz = B.gz(i, j, k, 2);
      const double wx = B.gx(i, j, k, 3), wy = B.gy(i, j, k, 3), wz = B.gz(i, j, k, 3);
      const double vort = std::sqrt(((wy - vz) * (wy - vz) + (uz - wx) * (uz - wx) + (vx - uy) * (vx - uy)));
      const double strain = std::sqrt(((wy + vz) * (wy + vz) + (uz + wx) * (uz + wx) + (vx + uy) * (vx + uy) + 2 * (ux * ux) + 2 * (vy * vy) + 2 * (wz * wz)));
      const double TuL = std::fmin(100.0 * std::sqrt(2.0 * B.qp(i, j, k, 6) / 3.0) / (B.qp(i, j, k, 7) * d), 100.0);
      const double Re_theta = 100.0 + 1000.0 * std::exp(-TuL);
      const double Rev = density * d * d * strain / muc;
      const double RT = density * B.qp(i, j, k, 6) / (muc * B.qp(i, j, k, 7));
      const double hr = 0.5 * RT;
      const double Fturb = std::exp(-((hr * hr) * (hr * hr)));
      const double Fonset1 = Rev / (2.2 * Re_theta);
      const double Fonset2 = std::fmin(Fonset1, 2.0);
      const double r35 = RT / 3.5;
      const double Fonset3 = std::fmax(1.0 - (r35 * r35 * r35), 0.0);
      const double Fonset = std::fmax(Fonset2 - Fonset3, 0.0);
      const double Dp = 100 * density * strain * Fonset * (1.0 - 2.0 * Q[0][7]);
      const double De = 0.06 * vort * Fturb * density * (2.0 * 50.0 * Q[0][7] - 1.0);
      D[7] = (D[7] + (-Dp + De) * B.cells.vol(i, j, k));
    }
    if (MODEL == M_KKL) {   // lusgs.f90:1339-1341
      const double vol = B.cells.vol(i, j, k), d = B.dist(i, j, k);
      D[5] = D[5] + (2.5 * (std::pow(kkl_cmu_l, (0.75))) * Q[0][0] * (std::pow(Q[0][5], (1.5))) * vol / Q[0][6]);
      D[5] = D[5] + (2 * mmu(i, j, k) * vol / (d * d));
      D[6] = D[6] + (6 * mmu(i, j, k) * vol / (d * d));
    }
    if (MODEL == M_SA) {   // lusgs.f90:1868-1913: the source-term derivatives go into EVERY component of D (array assignment) -- reproduced
      const double vol = B.cells.vol(i, j, k);
      const double density = B.qp(i, j, k, 1);
      const double a = (B.gy(i, j, k, 3) - B.gz(i, j, k, 2)), b = (B.gz(i, j, k, 1) - B.gx(i, j, k, 3)), cc = (B.gx(i, j, k, 2) - B.gy(i, j, k, 1));
      const double Omega = std::sqrt(((a * a) + (b * b) + (cc * cc)));
      const double dist_i = B.dist(i, j, k), dist_i_2 = dist_i * dist_i, k2 = sa_kappa * sa_kappa;
      const double nu = B.mu(i, j, k) / density;
      const double Ji = Q[0][5] / nu, Ji_2 = Ji * Ji, Ji_3 = Ji_2 * Ji;
      const double fv1 = (Ji_3) / ((Ji_3) + (sa_cv1_3));
      const double fv2 = 1.0 - Ji / (1.0 + (Ji * fv1));
      const double inv_k2_d2 = 1.0 / (k2 * dist_i_2);
      double Shat = Omega + Q[0][5] * fv2 * inv_k2_d2;
      Shat = std::fmax(Shat, 1.0e-10);
      const double inv_Shat = 1.0 / Shat;
      const double den1 = (Ji_3 + sa_cv1_3);
      const double dfv1 = 3.0 * Ji_2 * sa_cv1_3 / (nu * (den1 * den1));
      const double den2 = (1.0 + Ji * fv1);
      const double dfv2 = -((1.0 / nu) - Ji_2 * dfv1) / (den2 * den2);
      const double dShat = (fv2 + Q[0][5] * dfv2) * inv_k2_d2;
      const double r = std::fmin(Q[0][5] * inv_Shat * inv_k2_d2, 10.0);
      const double r2 = r * r, r6 = r2 * r2 * r2;
      const double g = r + sa_cw2 * ((r6) - r);
      const double g2 = g * g, g_6 = g2 * g2 * g2;
      const double glim = std::pow((1.0 + sa_cw3_6) / (g_6 + sa_cw3_6), (1.0 / 6.0));
      const double fw = g * glim;
      const double dr = (Shat - Q[0][5] * dShat) * inv_Shat * inv_Shat * inv_k2_d2;
      const double dg = dr * (1.0 + sa_cw2 * (6.0 * (r2 * r2 * r) - 1.0));
      const double dfw = dg * glim * (1.0 - g_6 / (g_6 + sa_cw3_6));
      for (int l = 0; l < NV; ++l) {
        D[l] = D[l] - sa_cb1 * (Q[0][5] * dShat + Shat) * vol;
        D[l] = D[l] + sa_cw1 * (dfw * Q[0][5] + 2 * fw) * Q[0][5] / dist_i_2 * vol;
      }
    }
  };
  const double zero[NV] = {0};
  // forward sweep
  for (int k = 1; k <= kmx - 1; ++k)
    for (int j = 1; j <= jmx - 1; ++j)
      for (int i = 1; i <= imx - 1; ++i) {
        double Q[7][NV], L[6], D[NV]; Face F[6];
        cell_setup(i, j, k, Q, F, L, D);
        double Del[3][NV], DQ[3][NV];
        for (int n = 0; n < 3; ++n) {
          for (int l = 0; l < NV; ++l) DQ[n][l] = delQstar(i + di[n], j + dj[n], k + dk[n], l + 1);
          double Fn[NV], Fo[NV];
          lusgs_flux<NV, MODEL>(X, Q[n + 1], Q[0], DQ[n], F[n], Fn);
          lusgs_flux<NV, MODEL>(X, Q[n + 1], Q[0], zero, F[n], Fo);
          for (int l = 0; l < NV; ++l) Del[n][l] = Fn[l] - Fo[l];
        }
        for (int l = 0; l < NV; ++l) {
          const double deltaU = -B.residue(i, j, k, l + 1) -
                                0.5 * (((Del[0][l] - L[0] * DQ[0][l]) + (Del[1][l] - L[1] * DQ[1][l])) + (Del[2][l] - L[2] * DQ[2][l]));
          delQstar(i, j, k, l + 1) = deltaU / D[l];
        }
      }
  // backward sweep
  for (int i = imx - 1; i >= 1; --i)
    for (int j = jmx - 1; j >= 1; --j)
      for (int k = kmx - 1; k >= 1; --k) {
        double Q[7][NV], L[6], D[NV]; Face F[6];
        cell_setup(i, j, k, Q, F, L, D);
        double Del[3][NV], DQ[3][NV];
        for (int n = 0; n < 3; ++n) {
          for (int l = 0; l < NV; ++l) DQ[n][l] = delQ(i + di[n + 3], j + dj[n + 3], k + dk[n + 3], l + 1);
          double Fn[NV], Fo[NV];
          lusgs_flux<NV, MODEL>(X, Q[n + 4], Q[0], DQ[n], F[n + 3], Fn);
          lusgs_flux<NV, MODEL>(X, Q[n + 4], Q[0], zero, F[n + 3], Fo);
          for (int l = 0; l < NV; ++l) Del[n][l] = Fn[l] - Fo[l];
        }
        for (int l = 0; l < NV; ++l)
          delQ(i, j, k, l + 1) = delQstar(i, j, k, l + 1) -
                                 0.5 * (((Del[0][l] - L[3] * DQ[0][l]) + (Del[1][l] - L[4] * DQ[1][l])) + (Del[2][l] - L[5] * DQ[2][l])) / D[l];
      }
  // conservative update (no positivity check in the reference: a negative density or pressure surfaces as NaN in the next residual)
  for (int k = 1; k <= kmx - 1; ++k)
    for (int j = 1; j <= jmx - 1; ++j)
      for (int i = 1; i <= imx - 1; ++i) {
        double cq[NV];
        cq[0] = B.qp(i, j, k, 1);
        cq[1] = B.qp(i, j, k, 1) * B.qp(i, j, k, 2);
        cq[2] = B.qp(i, j, k, 1) * B.qp(i, j, k, 3);
        cq[3] = B.qp(i, j, k, 1) * B.qp(i, j, k, 4);
        cq[4] = (B.qp(i, j, k, 5) / (c.gm - 1.0)) +
                (0.5 * B.qp(i, j, k, 1) * (((B.qp(i, j, k, 2) * B.qp(i, j, k, 2)) + (B.qp(i, j, k, 3) * B.qp(i, j, k, 3))) + (B.qp(i, j, k, 4) * B.qp(i, j, k, 4))));
        if (NV >= 6) cq[5] = B.qp(i, j, k, 1) * B.qp(i, j, k, 6);
        if (NV >= 7) cq[6] = B.qp(i, j, k, 1) * B.qp(i, j, k, 7);
        if (NV == 8) cq[7] = B.qp(i, j, k, 1) * B.qp(i, j, k, 8);
        for (int l = 0; l < NV; ++l) cq[l] = cq[l] + delQ(i, j, k, l + 1);
        B.qp(i, j, k, 1) = cq[0];
        B.qp(i, j, k, 2) = cq[1] / cq[0];
        B.qp(i, j, k, 3) = cq[2] / cq[0];
        B.qp(i, j, k, 4) = cq[3] / cq[0];
        B.qp(i, j, k, 5) = (c.gm - 1.0) * (cq[4] - (0.5 * (((cq[1] * cq[1]) + (cq[2] * cq[2])) + (cq[3] * cq[3])) / cq[0]));
        if (MODEL == M_SST || MODEL == M_LCTM) {
          if (cq[5] > 0) B.qp(i, j, k, 6) = cq[5] / cq[0];
          if (cq[6] > 0) B.qp(i, j, k, 7) = cq[6] / cq[0];
        }
        if (MODEL == M_LCTM) B.qp(i, j, k, 8) = std::fmax(cq[7] / cq[0], 0.0);   // lusgs.f90:2663-2664: here the intermittency IS advanced
        if (MODEL == M_KKL) {   // lusgs.f90:1505-1508
          B.qp(i, j, k, 6) = std::fmax(cq[5] / cq[0], 1.e-8);
          B.qp(i, j, k, 7) = std::fmax(cq[6] / cq[0], 1.e-8);
        }
        if (MODEL == M_SA) B.qp(i, j, k, 6) = std::fmax(cq[5] / cq[0], 1.e-8);   // lusgs.f90:2095-2096
      }
}

}  // namespace

// lusgs.f90:134-183: laminar / inviscid, sst / sst2003 (transition none | bc | lctm2015), kkl, sa (the dispatcher ignores transition = bc
// for sa): every routine of the dispatcher
int Block::update_with_lusgs() {
  if (c.turbulence == ORC_TURB_NONE) { lusgs_update<5, M_LAM>(*this); return 0; }
  if ((c.turbulence == ORC_TURB_SST || c.turbulence == ORC_TURB_SST2003) && c.transition != 2) { lusgs_update<7, M_SST>(*this); return 0; }
  if ((c.turbulence == ORC_TURB_SST || c.turbulence == ORC_TURB_SST2003) && c.transition == 2) { lusgs_update<8, M_LCTM>(*this); return 0; }
  if (c.turbulence == ORC_TURB_KKL) { lusgs_update<7, M_KKL>(*this); return 0; }
  if (c.turbulence == ORC_TURB_SA) { lusgs_update<6, M_SA>(*this); return 0; }
  return 64;
}

}  // namespace orc
