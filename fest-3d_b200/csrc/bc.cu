// Ghost-cell boundary fills (small kernels, O(N^2/3) work): one thread per boundary-face cell runs the reference's
// per-face sequence for that cell (three ghost layers, all variables, in the reference's statement order), then the
// edge / corner fill.  Faces are processed in the reference order imin,imax,jmin,jmax,kmin,kmax on one stream because
// far-field / periodic faces copy whole planes that include cells written by earlier faces.
//
// Reference: src/boundary/bc_primitive.f90:55-226 (dispatch + edge/corner), :229-562 (inlet/outlet/wall/slip/pole/fix),
// :564-643 (set_omega_at_wall), :645-1234 (far_field), :1778-1946 (temp_based_density), :1948-1978 (periodic);
// src/boundary/copy_bc.f90:58-131 (copy3); src/boundary/FT_bc.f90:15-107 (flow_tangency, incl. its Ifaces-normal defect).
#include "ctx.hpp"
#include "physics.cuh"
#include <algorithm>

namespace f3d {

struct FaceFrame {
  int face;        // 1..6
  long long sn;    // stride along the face-normal axis, signed so that +sn points INTO the domain
  long long sa, sb;
  long long c1;    // index of the first interior cell at (a,b) = (1,1)
  long long f1;    // index of the boundary face record at (a,b) = (1,1) in the face-direction arrays
  int na, nb;
};

__host__ __device__ inline FaceFrame make_frame(const Layout& L, int face) {
  FaceFrame f;
  f.face = face;
  const int ax = (face - 1) / 2;
  const bool lo = (face % 2) == 1;
  const long long st[3] = {1, L.sj, L.sk};
  const int mx[3] = {L.imx, L.jmx, L.kmx};
  const int a_ax = (ax == 0) ? 1 : 0, b_ax = (ax == 2) ? 1 : 2;
  f.sa = st[a_ax]; f.sb = st[b_ax];
  f.na = mx[a_ax] - 1; f.nb = mx[b_ax] - 1;
  f.sn = lo ? st[ax] : -st[ax];
  int idx[3] = {1, 1, 1};
  idx[ax] = lo ? 1 : mx[ax] - 1;
  f.c1 = L.idx(idx[0], idx[1], idx[2]);
  idx[ax] = lo ? 1 : mx[ax];
  f.f1 = L.idx(idx[0], idx[1], idx[2]);
  return f;
}

// interior layer l (1..3) and ghost layer l of the cell column at offset o = (a-1)*sa + (b-1)*sb
#define INT_(l) (fr.c1 + o + (long long)((l) - 1) * fr.sn)
#define GHO_(l) (fr.c1 + o - (long long)(l) * fr.sn)

__device__ __forceinline__ void copy3_flat(double* __restrict__ v, const FaceFrame& fr, long long o) {
  const double x = v[INT_(1)];
  v[GHO_(1)] = (1. * x - 0. * v[INT_(2)]) / 1.;
  v[GHO_(2)] = (1. * x - 0. * v[INT_(2)]) / 1.;
  v[GHO_(3)] = (1. * x - 0. * v[INT_(2)]) / 1.;
}
__device__ __forceinline__ void copy3_anti(double* __restrict__ v, const FaceFrame& fr, long long o) {
#pragma unroll
  for (int l = 1; l <= 3; ++l) v[GHO_(l)] = (-1. * v[INT_(l)] - 0. * v[INT_(l + 1)]) / 1.;
}
__device__ __forceinline__ void copy3_symm(double* __restrict__ v, const FaceFrame& fr, long long o, double c1, double c2, double c3) {
#pragma unroll
  for (int l = 1; l <= 3; ++l) v[GHO_(l)] = (c2 * v[INT_(l)] - c3 * v[INT_(l + 1)]) / c1;
}
__device__ __forceinline__ void fix3(double* __restrict__ v, const FaceFrame& fr, long long o, double val) {
  v[GHO_(1)] = val; v[GHO_(2)] = val; v[GHO_(3)] = val;
}

// Riemann-invariant normal velocity at the boundary for the cell column at offset o (bc_primitive.f90:676-689)
__device__ __forceinline__ void far_field_state(const Params& P, const double* __restrict__ q, const double* __restrict__ gn,
                                                const FaceFrame& fr, long long o, double& Unb, double& Cb, double& Unexp, double& Uninf,
                                                double& nx, double& ny, double& nz) {
  const long long fs = P.L.fs;
  const long long c = INT_(1), f = fr.f1 + o;
  const double sg = (fr.face % 2 == 1) ? -1.0 : 1.0;
  nx = sg * gn[fs + f]; ny = sg * gn[2 * fs + f]; nz = sg * gn[3 * fs + f];
  const double cexp = sqrt(P.gm * q[4 * fs + c] / q[c]);
  const double cinf = sqrt(P.gm * P.pressure_inf / P.density_inf);
  Unexp = q[fs + c] * nx + q[2 * fs + c] * ny + q[3 * fs + c] * nz;
  Uninf = P.x_speed_inf * nx + P.y_speed_inf * ny + P.z_speed_inf * nz;
  const double Rinf = Uninf - 2 * cinf / (P.gm - 1.);
  const double Rexp = Unexp + 2 * cexp / (P.gm - 1.);
  Unb = 0.5 * (Rexp + Rinf);
  Cb = 0.25 * (P.gm - 1.) * (Rexp - Rinf);
}

// face_or_mask > 0: that face; < 0: -mask of faces processed together (blockIdx.z = face-1).  Faces whose fill only reads
// interior cells and only writes their own ghost cells (every id except far-field -8 and periodic -9 / -10) are independent
// of each other and go in one launch.
__global__ void k_bc_face(const Params P, double* __restrict__ q, const double* __restrict__ geom, int face_or_mask) {
  int face = face_or_mask;
  if (face_or_mask < 0) {
    face = blockIdx.z + 1;
    if (!((-face_or_mask) >> (face - 1) & 1)) return;
  }
  const FaceFrame fr = make_frame(P.L, face);
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y * blockDim.y + threadIdx.y;
  if (a >= fr.na || b >= fr.nb) return;
  const long long o = (long long)a * fr.sa + (long long)b * fr.sb;
  const long long fs = P.L.fs;
  const int id = P.bc_id[face - 1];
  const bool sst = P.sst != 0;   // two-equation layout (sst, sst2003, kkl)
  const bool sa = P.sa != 0;   // the SA variable (field 5) follows the pattern of tk on every face (bc_primitive.f90:246 ... 543)
  // k-kL: kL (field 6) follows the pattern of omega with its own fixed value (fixed_tkl) -- except at the subsonic inlet, which fixes it
  // to fixed_tw (bc_primitive.f90:333-336, reproduced), and at the wall, which takes the anti copy instead of the wall-omega rule (:548-550)
  const bool kkl = P.kkl != 0;
  const int FIX_7 = kkl ? F3D_FIX_TKL : F3D_FIX_TW;
  double* rho = q; double* u = q + fs; double* v = q + 2 * fs; double* w = q + 3 * fs; double* p = q + 4 * fs;
  double* tk = q + 5 * fs; double* tw = q + 6 * fs;
  // transition = lctm2015: the intermittency (field 7) is fixed at the two inlets and copied flat on every other face, the wall included
  // (bc_primitive.f90:259-265, 299-304, 340-346, 384-389, 436-441, 474-479, 554-559); far-field / total pressure: as k (last cell decides)
  const bool lctm = P.lctm != 0;
  double* tgm = q + 7 * fs;
  const double(*fx)[6] = P.fixed;
  const int fi = face - 1;
  const int ax = (face - 1) / 2;
  switch (id) {
    case -1:   // supersonic_inlet
      if (P.current_iter <= 2) {
        fix3(rho, fr, o, fx[F3D_FIX_DENSITY][fi]); fix3(u, fr, o, fx[F3D_FIX_X_SPEED][fi]); fix3(v, fr, o, fx[F3D_FIX_Y_SPEED][fi]);
        fix3(w, fr, o, fx[F3D_FIX_Z_SPEED][fi]); fix3(p, fr, o, fx[F3D_FIX_PRESSURE][fi]);
        if (sst) { fix3(tk, fr, o, fx[F3D_FIX_TK][fi]); fix3(tw, fr, o, fx[FIX_7][fi]); }
        if (sa) fix3(tk, fr, o, fx[F3D_FIX_TV][fi]);
        if (lctm) fix3(tgm, fr, o, fx[F3D_FIX_TGM][fi]);
      }
      break;
    case -2: case -7:   // supersonic_outlet, pole: everything flat
      copy3_flat(rho, fr, o); copy3_flat(u, fr, o); copy3_flat(v, fr, o); copy3_flat(w, fr, o); copy3_flat(p, fr, o);
      if (sst) { copy3_flat(tk, fr, o); copy3_flat(tw, fr, o); }
      if (sa) copy3_flat(tk, fr, o);
      if (lctm) copy3_flat(tgm, fr, o);
      break;
    case -3:   // subsonic_inlet
      if (P.current_iter <= 2) {
        fix3(rho, fr, o, fx[F3D_FIX_DENSITY][fi]); fix3(u, fr, o, fx[F3D_FIX_X_SPEED][fi]); fix3(v, fr, o, fx[F3D_FIX_Y_SPEED][fi]);
        fix3(w, fr, o, fx[F3D_FIX_Z_SPEED][fi]);
        if (sst) { fix3(tk, fr, o, fx[F3D_FIX_TK][fi]); fix3(tw, fr, o, fx[F3D_FIX_TW][fi]); }
        if (sa) fix3(tk, fr, o, fx[F3D_FIX_TV][fi]);
        if (lctm) fix3(tgm, fr, o, fx[F3D_FIX_TGM][fi]);
      }
      copy3_flat(p, fr, o);
      break;
    case -4:   // subsonic_outlet
      copy3_flat(rho, fr, o); copy3_flat(u, fr, o); copy3_flat(v, fr, o); copy3_flat(w, fr, o);
      if (P.current_iter <= 2) fix3(p, fr, o, fx[F3D_FIX_PRESSURE][fi]);
      if (sst) { copy3_flat(tk, fr, o); copy3_flat(tw, fr, o); }
      if (sa) copy3_flat(tk, fr, o);
      if (lctm) copy3_flat(tgm, fr, o);
      break;
    case -5: {  // wall: pressure symm, temp_based_density, no_slip (+ omega at wall)
      copy3_symm(p, fr, o, P.c1, P.c2, P.c3);
      const double T = fx[F3D_FIX_WALL_TEMP][fi];
      if (T < 0.0) {
        const long long c = INT_(1);
        const double stag = (p[c] / (P.R_gas * rho[c])) * (1 + (0.5 * (P.gm - 1.) * P.gm * p[c] / rho[c]));
#pragma unroll
        for (int l = 1; l <= 3; ++l) rho[GHO_(l)] = p[GHO_(l)] / (P.R_gas * stag);
      } else if (T > 1.0) {
#pragma unroll
        for (int l = 1; l <= 3; ++l) rho[GHO_(l)] = p[GHO_(l)] / (P.R_gas * (2 * T - (p[INT_(l)] / (P.R_gas * rho[INT_(l)]))));
      } else {
        copy3_symm(rho, fr, o, P.c1, P.c2, P.c3);
      }
      copy3_anti(u, fr, o); copy3_anti(v, fr, o); copy3_anti(w, fr, o);
      if (sa) copy3_anti(tk, fr, o);
      if (kkl) { copy3_anti(tk, fr, o); copy3_anti(tw, fr, o); }
      if (sst && !kkl) {
        copy3_anti(tk, fr, o);
        const double* dist = geom + (long long)G_DIST * fs;
        const long long g1 = GHO_(1), c = INT_(1);
        const double T_face = 0.5 * ((p[g1] / rho[g1]) + (p[c] / rho[c])) / P.R_gas;
        const double mu = P.mu_ref * pow(T_face / P.T_ref, 1.5) * ((P.T_ref + P.Sutherland_temp) / (T_face + P.Sutherland_temp));
        const double rh = 0.5 * (rho[g1] + rho[c]);
        const double d2 = 2 * dist[c];
#pragma unroll
        for (int l = 1; l <= 3; ++l) tw[GHO_(l)] = 120 * mu / (rh * kBeta1 * (d2 * d2)) - tw[INT_(l)];
      }
      if (lctm) copy3_flat(tgm, fr, o);
      break;
    }
    case -6: {  // slip_wall
      copy3_symm(rho, fr, o, P.c1, P.c2, P.c3); copy3_symm(p, fr, o, P.c1, P.c2, P.c3);
      if (sst) { copy3_symm(tk, fr, o, P.c1, P.c2, P.c3); copy3_symm(tw, fr, o, P.c1, P.c2, P.c3); }
      if (sa) copy3_symm(tk, fr, o, P.c1, P.c2, P.c3);
      if (lctm) copy3_flat(tgm, fr, o);
      // flow_tangency: dot with this direction's face normal, reflection with the I-face normal at the same index
      const double* gd = geom + (long long)(G_IA + 4 * ax) * fs;
      const double* gi = geom + (long long)G_IA * fs;
      const long long f = fr.f1 + o;
      const double dnx = gd[fs + f], dny = gd[2 * fs + f], dnz = gd[3 * fs + f];
      const double inx = gi[fs + f], iny = gi[2 * fs + f], inz = gi[3 * fs + f];
#pragma unroll
      for (int l = 1; l <= 3; ++l) {
        const long long c = INT_(l), g = GHO_(l);
        const double dot = u[c] * dnx + v[c] * dny + w[c] * dnz;
        u[g] = u[c] - (2.0 * dot * inx);
        v[g] = v[c] - (2.0 * dot * iny);
        w[g] = w[c] - (2.0 * dot * inz);
      }
      break;
    }
    case -8: {  // far_field: first ghost layer here, layers 2,3 by the whole-plane copy kernel
      const double* gn = geom + (long long)(G_IA + 4 * ax) * fs;
      double Unb, Cb, Unexp, Uninf, nx, ny, nz;
      far_field_state(P, q, gn, fr, o, Unb, Cb, Unexp, Uninf, nx, ny, nz);
      const long long c = INT_(1), g = GHO_(1);
      if (Unb > 0.) {
        const double vd = Unb - Unexp;
        u[g] = u[c] + vd * nx; v[g] = v[c] + vd * ny; w[g] = w[c] + vd * nz;
        const double s = p[c] / pow(rho[c], P.gm);
        rho[g] = pow(Cb * Cb / (P.gm * s), 1. / (P.gm - 1.));
        p[g] = (rho[g] * Cb * Cb / P.gm);
      } else {
        const double vd = Unb - Uninf;
        u[g] = P.x_speed_inf + vd * nx; v[g] = P.y_speed_inf + vd * ny; w[g] = P.z_speed_inf + vd * nz;
        const double s = P.pressure_inf / pow(P.density_inf, P.gm);
        rho[g] = pow(Cb * Cb / (P.gm * s), 1. / (P.gm - 1.));
        p[g] = (rho[g] * Cb * Cb / P.gm);
      }
      if (sst || sa) {
        // The reference calls whole-face copy3("flat") / fix() from inside the per-cell loop (:700-757); the ghost k,omega
        // of the whole face therefore end up decided by the LAST cell of the loop: outflow there -> flat copy, else fixed.
        const long long olast = (long long)(fr.na - 1) * fr.sa + (long long)(fr.nb - 1) * fr.sb;
        double Ub2, Cb2, a2, b2, x2, y2, z2;
        far_field_state(P, q, gn, fr, olast, Ub2, Cb2, a2, b2, x2, y2, z2);
        if (Ub2 > 0.) { copy3_flat(tk, fr, o); if (sst) copy3_flat(tw, fr, o); if (lctm) copy3_flat(tgm, fr, o); }
        else if (sst) { fix3(tk, fr, o, fx[F3D_FIX_TK][fi]); fix3(tw, fr, o, fx[FIX_7][fi]); if (lctm) fix3(tgm, fr, o, fx[F3D_FIX_TGM][fi]); }
        else fix3(tk, fr, o, fx[F3D_FIX_TV][fi]);
      }
      break;
    }
    case -11: {  // total_pressure (bc_primitive.f90:1237-1776): far-field Riemann velocity, p from the fixed total pressure
      const double* gn = geom + (long long)(G_IA + 4 * ax) * fs;
      double Unb, Cb, Unexp, Uninf, nx, ny, nz;
      far_field_state(P, q, gn, fr, o, Unb, Cb, Unexp, Uninf, nx, ny, nz);
      const long long c = INT_(1), g = GHO_(1);
      if (Unb > 0.) {
        const double vd = Unb - Unexp;
        u[g] = u[c] + vd * nx; v[g] = v[c] + vd * ny; w[g] = w[c] + vd * nz;
      } else {
        const double vd = Unb - Uninf;
        u[g] = P.x_speed_inf + vd * nx; v[g] = P.y_speed_inf + vd * ny; w[g] = P.z_speed_inf + vd * nz;
      }
      const long long m = (face == 5) ? c : g;   // kmin takes Mb from the interior cell (:1678), the other faces from the ghost
      const double Mb = sqrt(u[m] * u[m] + v[m] * v[m] + w[m] * w[m]) / Cb;
      p[g] = fx[F3D_FIX_TPRESSURE][fi] / pow((1 + 0.5 * (P.gm - 1.) * Mb * Mb), P.gm / (P.gm - 1.));
      rho[g] = P.gm * p[g] / (Cb * Cb);
      if (sst || sa) {   // whole-face copy3("flat") / fix() from inside the per-cell loop: the LAST cell of the loop decides
        const long long olast = (long long)(fr.na - 1) * fr.sa + (long long)(fr.nb - 1) * fr.sb;
        double Ub2, Cb2, a2, b2, x2, y2, z2;
        far_field_state(P, q, gn, fr, olast, Ub2, Cb2, a2, b2, x2, y2, z2);
        if (Ub2 > 0.) { copy3_flat(tk, fr, o); if (sst) copy3_flat(tw, fr, o); if (lctm) copy3_flat(tgm, fr, o); }
        else if (sst) { fix3(tk, fr, o, fx[F3D_FIX_TK][fi]); fix3(tw, fr, o, fx[FIX_7][fi]); if (lctm) fix3(tgm, fr, o, fx[F3D_FIX_TGM][fi]); }
        else fix3(tk, fr, o, fx[F3D_FIX_TV][fi]);
      }
      break;
    }
    default: break;   // interface (>= 0), -10 (multi-block periodic), -9 handled by the slab kernel
  }
}

// far_field / total_pressure: qp(-1,:,:,:) = qp(0,:,:,:); qp(-2,:,:,:) = qp(0,:,:,:) over the WHOLE plane incl. ghost rows
// (:762-763, :1344-1345)
__global__ void k_plane_copy(const Params P, double* __restrict__ q, int face) {
  const Layout& L = P.L;
  const int ax = (face - 1) / 2;
  const bool lo = (face % 2) == 1;
  const int a_ax = (ax == 0) ? 1 : 0, b_ax = (ax == 2) ? 1 : 2;
  const int mx[3] = {L.imx, L.jmx, L.kmx};
  const long long st[3] = {1, L.sj, L.sk};
  const int a = -2 + blockIdx.x * blockDim.x + threadIdx.x, b = -2 + blockIdx.y * blockDim.y + threadIdx.y;
  if (a > mx[a_ax] + 2 || b > mx[b_ax] + 2) return;
  int idx[3]; idx[a_ax] = a; idx[b_ax] = b; idx[ax] = lo ? 0 : mx[ax];
  const long long g = L.idx(idx[0], idx[1], idx[2]);
  const long long sn = lo ? -st[ax] : st[ax];
  for (int v = 0; v < L.nv; ++v) {
    const double x = q[v * L.fs + g];
    q[v * L.fs + g + sn] = x;
    q[v * L.fs + g + 2 * sn] = x;
  }
}

// periodic_bc (id -9): whole slabs incl. ghost rows, e.g. qp(-2:0,:,:,:) = qp(imx-3:imx-1,:,:,:)  (:1948-1978)
__global__ void k_periodic(const Params P, double* __restrict__ q, int face) {
  const Layout& L = P.L;
  const int ax = (face - 1) / 2;
  const bool lo = (face % 2) == 1;
  const int a_ax = (ax == 0) ? 1 : 0, b_ax = (ax == 2) ? 1 : 2;
  const int mx[3] = {L.imx, L.jmx, L.kmx};
  const int a = -2 + blockIdx.x * blockDim.x + threadIdx.x, b = -2 + blockIdx.y * blockDim.y + threadIdx.y;
  if (a > mx[a_ax] + 2 || b > mx[b_ax] + 2) return;
  for (int l = 0; l < 3; ++l) {
    int di[3], si[3];
    di[a_ax] = a; di[b_ax] = b; si[a_ax] = a; si[b_ax] = b;
    if (lo) { di[ax] = -2 + l; si[ax] = mx[ax] - 3 + l; } else { di[ax] = mx[ax] + l; si[ax] = 1 + l; }
    const long long d = L.idx(di[0], di[1], di[2]), s = L.idx(si[0], si[1], si[2]);
    for (int v = 0; v < L.nv; ++v) q[v * L.fs + d] = q[v * L.fs + s];
  }
}

// edge and corner fill with the reference's factor 0.33 and statement order (bc_primitive.f90:209-224):
// phase 0: the four edges along i (whole i extent), phase 1: the four edges along k, phase 2: the eight corners.
__global__ void k_edges(const Params P, double* __restrict__ q, int phase) {
  const Layout& L = P.L;
  const int imx = L.imx, jmx = L.jmx, kmx = L.kmx;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int e = blockIdx.y;   // which statement of the phase
  for (int v = 0; v < L.nv; ++v) {
    double* __restrict__ x = q + (long long)v * L.fs;
#define Q(i, j, k) x[L.idx(i, j, k)]
    if (phase == 0) {
      const int i = -2 + t;
      if (i > imx + 2) return;
      if (e == 0) Q(i, 0, 0) = 0.33 * (Q(i, 1, 1) + Q(i, 0, 1) + Q(i, 1, 0));
      else if (e == 1) Q(i, 0, kmx) = 0.33 * (Q(i, 1, kmx - 1) + Q(i, 0, kmx - 1) + Q(i, 1, kmx));
      else if (e == 2) Q(i, jmx, 0) = 0.33 * (Q(i, jmx - 1, 1) + Q(i, jmx, 1) + Q(i, jmx - 1, 0));
      else Q(i, jmx, kmx) = 0.33 * (Q(i, jmx - 1, kmx - 1) + Q(i, jmx, kmx - 1) + Q(i, jmx - 1, kmx));
    } else if (phase == 1) {
      const int k = -2 + t;
      if (k > kmx + 2) return;
      if (e == 0) Q(imx, 0, k) = 0.33 * (Q(imx - 1, 1, k) + Q(imx - 1, 0, k) + Q(imx, 1, k));
      else if (e == 1) Q(0, 0, k) = 0.33 * (Q(1, 1, k) + Q(1, 0, k) + Q(0, 1, k));
      else if (e == 2) Q(0, jmx, k) = 0.33 * (Q(1, jmx - 1, k) + Q(1, jmx, k) + Q(0, jmx - 1, k));
      else Q(imx, jmx, k) = 0.33 * (Q(imx - 1, jmx - 1, k) + Q(imx - 1, jmx, k) + Q(imx, jmx - 1, k));
    } else {
      if (t != 0) return;
      switch (e) {
        case 0: Q(0, 0, 0) = 0.33 * (Q(1, 0, 0) + Q(0, 1, 0) + Q(0, 0, 1)); break;
        case 1: Q(imx, 0, 0) = 0.33 * (Q(imx - 1, 0, 0) + Q(imx, 1, 0) + Q(imx, 0, 1)); break;
        case 2: Q(0, jmx, 0) = 0.33 * (Q(1, jmx, 0) + Q(0, jmx - 1, 0) + Q(0, jmx, 1)); break;
        case 3: Q(0, 0, kmx) = 0.33 * (Q(1, 0, kmx) + Q(0, 1, kmx) + Q(0, 0, kmx - 1)); break;
        case 4: Q(imx, jmx, 0) = 0.33 * (Q(imx - 1, jmx, 0) + Q(imx, jmx - 1, 0) + Q(imx, jmx, 1)); break;
        case 5: Q(imx, 0, kmx) = 0.33 * (Q(imx - 1, 0, kmx) + Q(imx, 1, kmx) + Q(imx, 0, kmx - 1)); break;
        case 6: Q(0, jmx, kmx) = 0.33 * (Q(1, jmx, kmx) + Q(0, jmx - 1, kmx) + Q(0, jmx, kmx - 1)); break;
        default: Q(imx, jmx, kmx) = 0.33 * (Q(imx - 1, jmx, kmx) + Q(imx, jmx - 1, kmx) + Q(imx, jmx, kmx - 1)); break;
      }
    }
#undef Q
  }
}

int launch_bc(Ctx* ctx) {
  const Layout& L = ctx->P.L;
  const int mx[3] = {L.imx, L.jmx, L.kmx};
  // order-independent faces in one launch; far-field / periodic faces keep the reference order imin..kmax among themselves
  // and run after it (their whole-plane copies read ghost rows the other faces have filled: bc_primitive.f90:762-763, 1948-1978)
  int mask = 0, nmax_a = 1, nmax_b = 1;
  bool ordered = false;
  for (int face = 1; face <= 6; ++face) {
    const int id = ctx->P.bc_id[face - 1];
    if (id == -8 || id == -9 || id == -11) ordered = true;
  }
  // a fill reads interior layers 1..4 along its normal: with fewer than 4 cells there (quasi-2-D blocks, kmx == 2) those are the
  // ghost cells of the opposite face, so the reference order (low face first, then the high face reading its fresh ghosts) matters
  if (std::min(std::min(L.imx, L.jmx), L.kmx) - 1 < 4) ordered = true;
  for (int face = 1; face <= 6; ++face) {
    const int id = ctx->P.bc_id[face - 1];
    if (id >= 0 || id == -10 || ordered) continue;
    const int ax = (face - 1) / 2, a_ax = (ax == 0) ? 1 : 0, b_ax = (ax == 2) ? 1 : 2;
    mask |= 1 << (face - 1);
    nmax_a = std::max(nmax_a, mx[a_ax] - 1); nmax_b = std::max(nmax_b, mx[b_ax] - 1);
  }
  if (mask) {
    dim3 block(32, 4), grid((nmax_a + 31) / 32, (nmax_b + 3) / 4, 6);
    k_bc_face<<<grid, block, 0, ctx->stream>>>(ctx->P, ctx->qp, ctx->geom, -mask);
    ctx->launches++;
  }
  for (int face = 1; face <= 6 && ordered; ++face) {
    const int id = ctx->P.bc_id[face - 1];
    if (id >= 0 || id == -10) continue;
    const int ax = (face - 1) / 2;
    const int a_ax = (ax == 0) ? 1 : 0, b_ax = (ax == 2) ? 1 : 2;
    dim3 block(32, 4);
    if (id == -9) {
      dim3 grid((mx[a_ax] + 5 + 31) / 32, (mx[b_ax] + 5 + 3) / 4);
      k_periodic<<<grid, block, 0, ctx->stream>>>(ctx->P, ctx->qp, face);
      ctx->launches++;
      continue;
    }
    dim3 grid((mx[a_ax] - 1 + 31) / 32, (mx[b_ax] - 1 + 3) / 4);
    k_bc_face<<<grid, block, 0, ctx->stream>>>(ctx->P, ctx->qp, ctx->geom, face);
    ctx->launches++;
    if (id == -8 || id == -11) {
      dim3 g2((mx[a_ax] + 5 + 31) / 32, (mx[b_ax] + 5 + 3) / 4);
      k_plane_copy<<<g2, block, 0, ctx->stream>>>(ctx->P, ctx->qp, face);
      ctx->launches++;
    }
  }
  k_edges<<<dim3((L.imx + 5 + 63) / 64, 4), 64, 0, ctx->stream>>>(ctx->P, ctx->qp, 0);
  k_edges<<<dim3((L.kmx + 5 + 63) / 64, 4), 64, 0, ctx->stream>>>(ctx->P, ctx->qp, 1);
  k_edges<<<dim3(1, 8), 32, 0, ctx->stream>>>(ctx->P, ctx->qp, 2);
  ctx->launches += 3;
  F3D_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace f3d
