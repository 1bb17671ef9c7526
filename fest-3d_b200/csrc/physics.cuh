// Per-face / per-cell device physics of the explicit path.  Everything is FP64 (reference: wp => real64,
// src/vartypes.f90:3).  Each function cites the reference routine whose arithmetic it reproduces; the code is
// organised for registers (state vectors as fixed-size arrays, switches as selects), not as array sweeps.
#pragma once
#include <cuda_runtime.h>
#include "../../include/fest3d_gpu.h"

namespace f3d {

// SST closure constants (src/global/global_sst.f90:6-16)
__device__ constexpr double kSigmaK1 = 0.85, kSigmaK2 = 1.0, kSigmaW1 = 0.5, kSigmaW2 = 0.856;
__device__ constexpr double kBeta1 = 0.075, kBeta2 = 0.0828, kBstar = 0.09, kA1 = 0.31;
// k-kL closure constants (src/global/global_kkl.f90:6-15); sigma_k = sigma_phi = 1
__device__ constexpr double kKklZeta1 = 1.2, kKklZeta2 = 0.97, kKklZeta3 = 0.13, kKklCmu = 0.09, kKklKappa = 0.41, kKklC11 = 10.0, kKklC12 = 1.3, kKklCd1 = 4.7;
__device__ constexpr double kKklCmu25 = 0.54772255750516607;   // cmu**0.25
__device__ constexpr double kKklCmu75 = 0.16431676725154984;   // cmu**0.75
// SA closure constants (src/global/global_sa.f90:6-19)
__device__ constexpr double kCb1 = 0.1355, kCb2 = 0.6220, kCw2 = 0.3, kCw3 = 2.0, kCv1 = 7.1, kSigmaSA = 2. / 3., kKappaSA = 0.41;
__device__ constexpr double kCw1 = (kCb1 / (kKappaSA * kKappaSA)) + ((1 + kCb2) / kSigmaSA);
__device__ __forceinline__ double pow3(double x) { return x * x * x; }
__device__ __forceinline__ double pow6(double x) { const double x2 = x * x; return x2 * x2 * x2; }
// gamma_BC of the algebraic Bas-Cakmakcioglu transition model (source.f90:570-585, 1156-1170)
// re_theta_t = 803.73 (Tu_inf + 0.6067)^-1.027 depends on the run's free-stream turbulence intensity only: evaluated once on the host
// (api.cu) instead of one pow per cell and stage
__device__ __forceinline__ double gamma_bc(double re_theta_t, double nu_cr, double nu_t, double vmag, double dist, double re_v) {
  const double chi_1 = 0.002;
  const double nu_bc = nu_t / (vmag * dist);
  const double re_theta = re_v / 2.193;
  const double term1 = sqrt(fmax(re_theta - re_theta_t, 0.) / (chi_1 * re_theta_t));
  const double term2 = sqrt(fmax(nu_bc - nu_cr, 0.0) / nu_cr);
  return 1.0 - exp(-(term1 + term2));
}
// SA wall function fw = g*((1+cw3^6)/(g^6+cw3^6))^(1/6) with g = r + cw2 (r^6 - r)   (source.f90:958-960, update.f90:416-418)
__device__ __forceinline__ double sa_fw(double r) {
  const double g = r + kCw2 * (pow6(r) - r);
  return g * cbrt(sqrt((1.0 + pow6(kCw3)) / (pow6(g) + pow6(kCw3))));   // x**(1/6) as a cube root of a square root: <= 2 ulp from pow, a third of its instructions
}

__device__ __forceinline__ double sgn1(double x) { return copysign(1.0, x); }          // sign(1.0, x)
__device__ __forceinline__ double sq(double x) { return x * x; }

// FP64 reciprocal / reciprocal square root without the IEEE slow path: MUFU seed (>= 20 good bits, relative error e with
// |e| < 2^-20) and ONE third-order correction r0*(1 + e + e^2) (error e^3 < 2^-60, i.e. < 1 ulp before the final rounding)
// -- 3 dependent DFMA instead of the 4 of two Newton steps, and a shorter chain.  An IEEE divide costs 14.4 DFMA issue
// slots on B200 (profiles/r01_fp64_ops_microbench.txt); results differ from a/b by <= ~1.5 ulp, five orders below the
// 1e-12 parity tolerance.  Operands on this path are finite and far from the subnormal / overflow range; the places where
// the reference relies on IEEE 0/0 or x/0 semantics (boundary_cell_face_values) keep the IEEE divide.
__device__ __forceinline__ double rcp64(double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  const double e = fma(-b, r, 1.0);
  const double t = fma(e, e, e);
  return fma(r, t, r);
}
__device__ __forceinline__ double fdiv(double a, double b) { return a * rcp64(b); }
// x > 0 only (rsqrt(0) = inf): y0*(1 + h/2 + 3h^2/8) with h = 1 - x*y0^2 (|h| < 2^-19 -> error < 2^-58)
__device__ __forceinline__ double rsqrt64(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = fma(-(x * y), y, 1.0);
  const double p = fma(0.375, h, 0.5) * h;
  return fma(y, p, y);
}
__device__ __forceinline__ double sqrt64(double x) {   // x > 0 only
  const double y = rsqrt64(x);
  const double s = x * y;
  return fma(fma(-s, s, x), 0.5 * y, s);
}
// min / max as compare + select (no NaN operands on these call sites)
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }
// non-linear weights of WENO-NM (weno_NM.f90:96-98): as ratios, multiplied through by the product of the three denominators
#ifdef F3D_WENO_IEEE
#define WENO_NM_WEIGHTS const double w1 = wdiv(0.1, sq(eps + B1)), w2 = wdiv(0.6, sq(eps + B2)), w3 = wdiv(0.3, sq(eps + B3));
#else
#define WENO_NM_WEIGHTS const double b1 = sq(eps + B1), b2 = sq(eps + B2), b3 = sq(eps + B3); \
                        const double w1 = 0.1 * (b2 * b3), w2 = 0.6 * (b1 * b3), w3 = 0.3 * (b1 * b2);
#endif
// divisions of the WENO / WENO-NM / PPM reconstructions (positive, well-scaled divisors; constant divisors 3, 6, 12)
#ifdef F3D_WENO_IEEE
__device__ __forceinline__ double wdiv(double a, double b) { return a / b; }
__device__ __forceinline__ double wthird(double a) { return a / 3.0; }
__device__ __forceinline__ double wsixth(double a) { return a / 6.0; }
__device__ __forceinline__ double wtwelfth(double a) { return a / 12.; }
#else
__device__ __forceinline__ double wdiv(double a, double b) { return a * rcp64(b); }
__device__ __forceinline__ double wthird(double a) { return a * (1.0 / 3.0); }
__device__ __forceinline__ double wsixth(double a) { return a * (1.0 / 6.0); }
__device__ __forceinline__ double wtwelfth(double a) { return a * (1.0 / 12.0); }
#endif
// max(0, 1 - floor(abs(M))) of the Mach splittings: 1 inside |M| < 1, else 0
__device__ __forceinline__ double subsonic(double M) { return fabs(M) < 1.0 ? 1.0 : 0.0; }

// ------------------------------------------------------------------------------------------------------------
// Face reconstruction: given the line of cell values around cell c, return the value this cell contributes to
// its high face (the "left" state there) and to its low face (the "right" state there).
//   muscl.f90:161-196 (Koren limiter, kappa = 1/3); weno.f90:55-91; weno_NM.f90:66-118; ppm.f90:44-105;
//   face_interpolant.f90:61-77 (first order).  q[0..6] = cells c-3 .. c+3 (only the used part need be valid).
template <int INTERP>
__device__ __forceinline__ void cell_face_values(const double* q, const double* vol, int limiter, double& to_hi, double& to_lo) {
  const double qm2 = q[1], qm1 = q[2], q0 = q[3], qp1 = q[4], qp2 = q[5];
  if (INTERP == F3D_INTERP_NONE) {
    to_hi = q0; to_lo = q0;
  } else if (INTERP == F3D_MUSCL) {
    const double fd = qp1 - q0, bd = q0 - qm1;
    const double kappa = 1. / 3.;
    if (limiter == 0) {   // psi = 1 - (1 - psi)*0 = 1 exactly
      to_hi = q0 + 0.25 * (((1. - kappa) * bd) + ((1. + kappa) * fd));
      to_lo = q0 - 0.25 * (((1. + kappa) * bd) + ((1. - kappa) * fd));
    } else {
      double r = fd * rcp64(bd + copysign(1e-14, bd));
      double psi1 = dmax(0., dmin(dmin(2 * r, (2. / 3.) * (r - 1.0) + 1.0), 2.));
      r = bd * rcp64(fd + copysign(1e-14, fd));
      double psi2 = dmax(0., dmin(dmin(2 * r, (2. / 3.) * (r - 1.0) + 1.0), 2.));
      psi1 = (1 - (1 - psi1) * limiter);
      psi2 = (1 - (1 - psi2) * limiter);
      to_hi = q0 + 0.25 * (((1. - kappa) * psi1 * bd) + ((1. + kappa) * psi2 * fd));
      to_lo = q0 - 0.25 * (((1. + kappa) * psi1 * bd) + ((1. - kappa) * psi2 * fd));
    }
  } else if (INTERP == F3D_WENO) {
    const double eps = 1e-6;
    double t, s;
    t = (qm2 - 2.0 * qm1 + q0); s = (qm2 - 4.0 * qm1 + 3.0 * q0);
    const double B1 = (13.0 / 12.0) * (t * t) + (1.0 / 4.0) * (s * s);
    t = (qm1 - 2.0 * q0 + qp1); s = (qm1 - qp1);
    const double B2 = (13.0 / 12.0) * (t * t) + (1.0 / 4.0) * (s * s);
    t = (q0 - 2.0 * qp1 + qp2); s = (3.0 * q0 - 4.0 * qp1 + qp2);
    const double B3 = (13.0 / 12.0) * (t * t) + (1.0 / 4.0) * (s * s);
    // divisions: eps + B >= 1e-6 and the weight sums are positive, so the fast reciprocal applies (<= 2 ulp, like every other
    // division of the sweep); x / 6 as x * (1/6).  The IEEE sequences cost 14.4 DFMA slots each, 11 per variable and direction:
    // profiles/r01_fp64_ops_microbench.txt.  F3D_WENO_IEEE keeps them.
#ifdef F3D_WENO_IEEE
    const double i1 = wdiv(1.0, sq(eps + B1)), i2 = wdiv(1.0, sq(eps + B2)), i3 = wdiv(1.0, sq(eps + B3));
#else
    // the weights g_i / (eps + B_i)^2 only enter as ratios: both sums are multiplied through by the product of the three
    // denominators, i.e. i_1 = b_2 b_3, i_2 = b_1 b_3, i_3 = b_1 b_2 with b_i = (eps + B_i)^2 -- three products instead of three
    // reciprocals, one normalising reciprocal per face value left (b_i between 1e-12 and ~1e20 for any physical field: no range issue)
    const double b1 = sq(eps + B1), b2 = sq(eps + B2), b3 = sq(eps + B3);
    const double i1 = b2 * b3, i2 = b1 * b3, i3 = b1 * b2;
#endif
    {
      const double P1 = wsixth(2.0 * qm2 - 7.0 * qm1 + 11.0 * q0);
      const double P2 = wsixth(-1.0 * qm1 + 5.0 * q0 + 2.0 * qp1);
      const double P3 = wsixth(2.0 * q0 + 5.0 * qp1 - 1.0 * qp2);
      const double w1 = 0.1 * i1, w2 = 0.6 * i2, w3 = 0.3 * i3;
      to_hi = wdiv((w1 * P1 + w2 * P2) + w3 * P3, (w1 + w2) + w3);
    }
    {  // the low-face value reuses the same smoothness indicators with mirrored linear weights (weno.f90:84-86)
      const double P1 = wsixth(2.0 * qp2 - 7.0 * qp1 + 11.0 * q0);
      const double P2 = wsixth(-1.0 * qp1 + 5.0 * q0 + 2.0 * qm1);
      const double P3 = wsixth(2.0 * q0 + 5.0 * qm1 - 1.0 * qm2);
      const double w1 = 0.1 * i3, w2 = 0.6 * i2, w3 = 0.3 * i1;
      to_lo = wdiv((w1 * P1 + w2 * P2) + w3 * P3, (w1 + w2) + w3);
    }
  } else if (INTERP == F3D_WENO_NM) {
    const double eps = 1e-6;
    const double vm2 = vol[1], vm1 = vol[2], v0 = vol[3], vp1 = vol[4], vp2 = vol[5];
    const double alpha12 = wdiv(vp2, vp1 + vp2), alpha01 = wdiv(vp1, v0 + vp1);   // volumes are positive (geometry.f90:476-494)
    const double alpha10 = wdiv(v0, vm1 + v0), alpha21 = wdiv(vm1, vm2 + vm1);
    const double U01 = (1.0 - alpha01) * q0 + alpha01 * qp1;
    const double U12 = (1.0 - alpha12) * qp1 + alpha12 * qp2;
    const double U10 = (1.0 - alpha10) * qm1 + alpha10 * q0;
    const double U21 = (1.0 - alpha21) * qm2 + alpha21 * qm1;
    const double U00 = qm1 + (1.0 - alpha21) * (qm1 - qm2);
    const double U11 = qp1 + alpha12 * (qp1 - qp2);
    double t, s;
    {
      const double P1 = wthird(6.0 * q0 - 1.0 * U10 - 2.0 * U00);
      const double P2 = wthird(-1.0 * U10 + 2.0 * q0 + 2.0 * U01);
      const double P3 = wthird(2.0 * U01 + 2.0 * qp1 - 1.0 * U12);
      t = (2 * U10 - 2.0 * U00); s = (4 * q0 - 2.0 * U10 - 2.0 * U00);
      const double B1 = (13.0 / 12.0) * (t * t) + (1.0 / 4.0) * (s * s);
      t = (2 * U10 - 4.0 * q0 + 2 * U01); s = (-2 * U10 + 2.0 * U01);
      const double B2 = (13.0 / 12.0) * (t * t) + (1.0 / 4.0) * (s * s);
      t = (2 * U01 - 4.0 * qp1 + 2 * U12); s = (-6 * U01 + 8.0 * qp1 - 2.0 * U12);
      const double B3 = (13.0 / 12.0) * (t * t) + (1.0 / 4.0) * (s * s);
      WENO_NM_WEIGHTS
      to_hi = wdiv((w1 * P1 + w2 * P2) + w3 * P3, (w1 + w2) + w3);
    }
    {
      const double P1 = wthird(6.0 * q0 - 1.0 * U01 - 2.0 * U11);
      const double P2 = wthird(-1.0 * U01 + 2.0 * q0 + 2.0 * U10);
      const double P3 = wthird(2.0 * U10 + 2.0 * qm1 - 1.0 * U21);
      t = (2 * U01 - 2.0 * U11); s = (4 * q0 - 2.0 * U01 - 2.0 * U11);
      const double B1 = (13.0 / 12.0) * (t * t) + (1.0 / 4.0) * (s * s);
      t = (2 * U01 - 4.0 * q0 + 2 * U10); s = (-2 * U01 + 2.0 * U10);
      const double B2 = (13.0 / 12.0) * (t * t) + (1.0 / 4.0) * (s * s);
      t = (2 * U10 - 4.0 * qm1 + 2 * U21); s = (-6 * U10 + 8.0 * qm1 - 2.0 * U21);
      const double B3 = (13.0 / 12.0) * (t * t) + (1.0 / 4.0) * (s * s);
      WENO_NM_WEIGHTS
      to_lo = wdiv((w1 * P1 + w2 * P2) + w3 * P3, (w1 + w2) + w3);
    }
  } else {  // PPM: 4-point face estimates on both faces of the cell, then the monotonicity fix of the cell
    double R = wtwelfth(7. * (q0 + qm1) - (qp1 + qm2));       // estimate at the low face
    double L = wtwelfth(7. * (qp1 + q0) - (qp2 + qm1));       // estimate at the high face
    if (limiter == 1) {
      if ((L - q0) * (q0 - R) <= 0) { L = q0; R = q0; }
      else {
        const double dqrl = L - R;
        const double dq6 = 6. * (q0 - 0.5 * (L + R));
        if (dqrl * dq6 > dqrl * dqrl) R = 3. * q0 - 2. * L;
        else if (-dqrl * dqrl > dqrl * dq6) L = 3. * q0 - 2. * R;
      }
    }
    to_hi = L; to_lo = R;
  }
}

// Boundary re-reconstruction of the first/last interior cell (boundary_state_reconstruction.f90:93-123):
// third-order MUSCL with the *unguarded* ratio fd/bd.  Fortran min/max with a NaN operand are taken with
// fmin/fmax semantics (0/0 on a uniform field picks the finite operand).
// The quotient of the unguarded ratios with the IEEE results the reference relies on for a vanishing divisor (x/0 = +-inf, 0/0 = NaN;
// a difference of two equal numbers is +0), and the fast reciprocal elsewhere (<= 2 ulp): no slow-path call in the kernel.
__device__ __forceinline__ double ratio_ieee0(double a, double b) {
  const double r = a * rcp64(b);
  const double z = (a == 0.0) ? __longlong_as_double(0x7ff8000000000000LL) : copysign(__longlong_as_double(0x7ff0000000000000LL), a);
  return (b == 0.0) ? z : r;
}
__device__ __forceinline__ void boundary_cell_face_values(double qm1, double q0, double qp1, int limiter, double& to_hi, double& to_lo) {
  const double fd = qp1 - q0, bd = q0 - qm1;
  double r = ratio_ieee0(fd, bd);
  double psi1 = fmax(0., fmin(fmin(2 * r, (2 + r) * (1. / 3.)), 2.));
  psi1 = (1 - (1 - psi1) * limiter);
  r = ratio_ieee0(bd, fd);
  double psi2 = fmax(0., fmin(fmin(2 * r, (2 + r) * (1. / 3.)), 2.));
  psi2 = (1 - (1 - psi2) * limiter);
  const double kappa = 1. / 3.;
  to_hi = q0 + 0.25 * (((1. - kappa) * psi1 * bd) + ((1. + kappa) * psi2 * fd));
  to_lo = q0 - 0.25 * (((1. + kappa) * psi1 * bd) + ((1. - kappa) * psi2 * fd));
}

// ------------------------------------------------------------------------------------------------------------
// Inviscid flux through one face, times the face area.  L/R = primitive (rho,u,v,w,p[,k,omega]).
//   van_leer.f90:56-146, ldfss0.f90:57-161, ausm.f90:56-154, ausmP.f90:77-198, ausmUP.f90:88-216, slau.f90:82-204.
// mask = make_{F,G,H}_flux_zero of the face (0 on wall / slip-wall / pole faces, bc.f90:53-66).
// Returns the face-averaged speed of sound 0.5*(c_L + c_R) that the local time step uses (time.f90:159-175).
// Divisions go through rcp64 (one reciprocal of each density and of the interface sound speed, reused).
template <int NV>
__device__ __forceinline__ double inviscid_flux(int scheme, double gm, double MInf, const double (&L)[NV], const double (&R)[NV],
                                                double A, double nx, double ny, double nz, double mask, bool flux_on, bool need_c,
                                                double (&F)[NV]) {
  const double g1 = gm / (gm - 1.);
  const double iL = rcp64(L[0]), iR = rcp64(R[0]);
  double cbar = 0.0;
  if (need_c || scheme <= F3D_AUSM || scheme == F3D_SLAU) cbar = 0.5 * (sqrt64(gm * L[4] * iL) + sqrt64(gm * R[4] * iR));
  if (!flux_on) {
#pragma unroll
    for (int l = 0; l < NV; ++l) F[l] = 0.0;
    return cbar;
  }
  const double VnL = L[1] * nx + L[2] * ny + L[3] * nz;
  const double VnR = R[1] * nx + R[2] * ny + R[3] * nz;
  const double HL = (0.5 * (L[1] * L[1] + L[2] * L[2] + L[3] * L[3])) + (g1 * L[4] * iL);
  const double HR = (0.5 * (R[1] * R[1] + R[2] * R[2] + R[3] * R[3])) + (g1 * R[4] * iR);
  if (scheme <= F3D_AUSM) {  // van Leer, LDFSS(0), AUSM: shared Mach / pressure splitting
    const double ic = rcp64(cbar);
    const double ML = VnL * ic, MR = VnR * ic;
    const double Mp = 0.25 * sq(1. + ML), Dp = 0.25 * sq(1. + ML) * (2. - ML);
    const double Mm = -0.25 * sq(1. - MR), Dm = 0.25 * sq(1. - MR) * (2. + MR);
#ifdef F3D_SPLIT_ARITH   // the reference's arithmetic form of the switches (van_leer.f90:79-101, ausm.f90:80-100)
    const double aP = 0.5 * (1.0 + sgn1(ML)), bL = -subsonic(ML);
    double cP = (aP * (1.0 + bL) * ML) - bL * Mp;
    const double sDp = (aP * (1. + bL)) - (bL * Dp);
    const double aM = 0.5 * (1.0 - sgn1(MR)), bR = -subsonic(MR);
    double cM = (aM * (1.0 + bR) * MR) - bR * Mm;
    const double sDm = (aM * (1. + bR)) - (bR * Dm);
#else
    // alpha = 0.5 (1 +- sign(1, M)) is 0 or 1 and beta = -max(0, 1 - int(|M|)) is 0 or -1, so the reference's blends
    // alpha (1 + beta) M - beta M+- and alpha (1 + beta) - beta D+- pick one of their terms.  Picked here by integer tests on the high
    // word of M (|M| < 1 <=> high word without its sign < that of 1.0, whose low word is zero; sign(1, -0.) = -1 like the
    // reference): the same values (a zero may differ in sign), 8 FP64 instructions fewer per side
    const int hML = __double2hiint(ML), hMR = __double2hiint(MR);
    const bool subL = (hML & 0x7fffffff) < 0x3ff00000, subR = (hMR & 0x7fffffff) < 0x3ff00000;
    const bool posL = hML >= 0, negR = hMR < 0;
    const double bL = subL ? -1.0 : 0.0, bR = subR ? -1.0 : 0.0;   // (LDFSS only)
    double cP = subL ? Mp : (posL ? ML : 0.0);
    const double sDp = subL ? Dp : (posL ? 1.0 : 0.0);
    double cM = subR ? Mm : (negR ? MR : 0.0);
    const double sDm = subR ? Dm : (negR ? 1.0 : 0.0);
#endif
    if (scheme == F3D_AUSM) {
      const double t = cP + cM;
      cP = dmax(0., t); cM = dmin(0., t);
    } else if (scheme == F3D_LDFSS0) {
      const double Ml = 0.25 * bL * bR * sq(sqrt((ML * ML + MR * MR) * 0.5) - 1);
      const double dp = L[4] - R[4];
      const double ic2 = ic * ic;
      cP = cP - Ml * (1 - dp * 0.5 * iL * ic2);
      cM = cM + Ml * (1 - dp * 0.5 * iR * ic2);
    }
    const double mP = (L[0] * cbar * cP) * mask;
    const double mM = (R[0] * cbar * cM) * mask;
    F[0] = mP * A + mM * A;
    F[1] = ((mP * L[1]) + (sDp * L[4] * nx)) * A + ((mM * R[1]) + (sDm * R[4] * nx)) * A;
    F[2] = ((mP * L[2]) + (sDp * L[4] * ny)) * A + ((mM * R[2]) + (sDm * R[4] * ny)) * A;
    F[3] = ((mP * L[3]) + (sDp * L[4] * nz)) * A + ((mM * R[3]) + (sDm * R[4] * nz)) * A;
    F[4] = (mP * HL) * A + (mM * HR) * A;
#pragma unroll
    for (int l = 5; l < NV; ++l) F[l] = (mP * L[l]) * A + (mM * R[l]) * A;
    return cbar;
  }
  double mass, pbar;
  if (scheme == F3D_SLAU) {
    const double C = cbar;
    const double iC = rcp64(C);
    const double ML = VnL * iC, MR = VnR * iC;
    const double sL = subsonic(ML), sR = subsonic(MR);
    const double bL = (1.0 - sL) * 0.5 * (1.0 + sgn1(ML)) + sL * 0.25 * (2.0 - ML) * sq(ML + 1.0);
    const double bR = (1.0 - sR) * 0.5 * (1.0 - sgn1(MR)) + sR * 0.25 * (2.0 + MR) * sq(MR - 1.0);
    const double vt = sqrt(0.5 * ((L[1] * L[1]) + (L[2] * L[2]) + (L[3] * L[3]) + (R[1] * R[1]) + (R[2] * R[2]) + (R[3] * R[3])));
    const double Xi = sq(1.0 - dmin(1.0, vt * iC));
    const double Vnabs = (L[0] * fabs(VnL) + R[0] * fabs(VnR)) * rcp64(L[0] + R[0]);
    const double fnG = -1.0 * dmax(dmin(ML, 0.0), -1.0) * dmin(dmax(MR, 0.0), 1.0);
    pbar = 0.5 * ((L[4] + R[4]) + (bL - bR) * (L[4] - R[4]) + (1.0 - Xi) * (bL + bR - 1.0) * (L[4] + R[4]));
    const double VaL = (1.0 - fnG) * Vnabs + fnG * fabs(VnL);
    const double VaR = (1.0 - fnG) * Vnabs + fnG * fabs(VnR);
    mass = 0.5 * ((L[0] * (VnL + VaL) + R[0] * (VnR - VaR)) - (Xi * (R[4] - L[4]) * iC));
  } else {  // AUSM+ and AUSM+-up
    const double cs2 = 2.0 * (gm - 1.0) * (0.5 * (HL + HR)) / (gm + 1.0);
    const double cs = sqrt64(cs2);
    const bool up = scheme == F3D_AUSMUP;
    const double cL = cs2 * rcp64(dmax(cs, up ? VnL : fabs(VnL)));
    const double cR = cs2 * rcp64(dmax(cs, up ? -VnR : fabs(VnR)));
    const double C = dmin(cL, cR);
    const double iC = rcp64(C);
    const double ML = VnL * iC, MR = VnR * iC;
    double alfa = 0.1875, fna = 1.0, Mb2 = 0.0;
    if (up) {
      Mb2 = 0.5 * ((VnL * VnL) + (VnR * VnR)) * (iC * iC);
      const double Mo = sqrt(dmin(1.0, dmax(Mb2, MInf * MInf)));
      fna = Mo * (2.0 - Mo);
      alfa = 3.0 * (-4.0 + (5.0 * fna * fna)) / 16.0;
    }
    const double sL = subsonic(ML), sR = subsonic(MR);
    double FmL = (0.5 * (1.0 + sgn1(ML)) * (1.0 - sL) * ML) + sL * 0.25 * sq(1.0 + ML);
    double bL = (0.5 * (1.0 + sgn1(ML)) * (1.0 - sL)) + sL * 0.25 * sq(1.0 + ML) * (2.0 - ML);
    double FmR = (0.5 * (1.0 - sgn1(MR)) * (1.0 - sR) * MR) - sR * 0.25 * sq(1.0 - MR);
    double bR = (0.5 * (1.0 - sgn1(MR)) * (1.0 - sR)) + sR * 0.25 * sq(1.0 - MR) * (2.0 + MR);
    const double tL = sq(ML * ML - 1.0), tR = sq(MR * MR - 1.0);
    FmL = FmL + sL * 0.125 * tL;
    bL = bL + sL * alfa * tL * ML;
    FmR = FmR - sR * 0.125 * tR;
    bR = bR - sR * alfa * tR * MR;
    double Mface = FmL + FmR;
    pbar = bL * L[4] + bR * R[4];
    if (up) {
      const double Pu = -0.75 * bL * bR * (L[0] + R[0]) * fna * C * (VnR - VnL);
      const double Mp = -2.0 * 0.25 * dmax(1.0 - (1.0 * Mb2), 0.0) * (R[4] - L[4]) * rcp64(fna * (L[0] + R[0]) * C * C);
      Mface = FmL + FmR + Mp;
      pbar = bL * L[4] + bR * R[4] + Pu;
    }
    mass = (Mface > 0.0) ? Mface * C * L[0] : Mface * C * R[0];
  }
  mass = mass * mask;
  const double mP = 0.5 * (mass + fabs(mass)), mM = 0.5 * (mass - fabs(mass));
  F[0] = (mP + mM) * A;
  F[1] = (((mP * L[1]) + (mM * R[1])) + (pbar * nx)) * A;
  F[2] = (((mP * L[2]) + (mM * R[2])) + (pbar * ny)) * A;
  F[3] = (((mP * L[3]) + (mM * R[3])) + (pbar * nz)) * A;
  F[4] = ((mP * HL) + (mM * HR)) * A;
#pragma unroll
  for (int l = 5; l < NV; ++l) F[l] = ((mP * L[l]) + (mM * R[l])) * A;
  return cbar;
}

}  // namespace f3d
