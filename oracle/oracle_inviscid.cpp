// TEST INFRASTRUCTURE ONLY -- CPU oracle, part 1: set-up, ghost-cell boundary conditions, face
// reconstruction, boundary-state override, inviscid fluxes, residue.  PARITY UNPINNED (see oracle_abi.h).
#include "oracle_core.hpp"

namespace orc {

static inline double fsign(double a, double b) { return std::copysign(std::fabs(a), b); }  // Fortran sign()
static inline double fmin3(double a, double b, double c) { return std::fmin(std::fmin(a, b), c); }

// ---------------------------------------------------------------------------------------------
// set-up: allocation shapes follow scheme.f90:52-63, face_interpolant.f90:43-55, gradients.f90:133-135,
// viscosity.f90:516-541, update.f90:85-125, bc.f90:48-66
void Block::setup(const OracleConfig& cfg) {
  c = cfg;
  imx = c.imx; jmx = c.jmx; kmx = c.kmx; nv = c.n_var;
  n_grad = (c.turbulence == ORC_TURB_SST || c.turbulence == ORC_TURB_SST2003 || c.turbulence == ORC_TURB_KKL) ? 6 : (c.turbulence == ORC_TURB_SA ? 5 : 4);   // gradients.f90:160-175
  if (c.transition == 2) n_grad += 1;   // lctm2015: the intermittency gradient is the last one (gradients.f90:259-264, :207-211)
  qp.alloc(-2, imx + 2, -2, jmx + 2, -2, kmx + 2, nv);
  Temp.alloc(-2, imx + 2, -2, jmx + 2, -2, kmx + 2);
  residue.alloc(1, imx - 1, 1, jmx - 1, 1, kmx - 1, nv);
  F.alloc(1, imx, 1, jmx - 1, 1, kmx - 1, nv);
  G.alloc(1, imx - 1, 1, jmx, 1, kmx - 1, nv);
  H.alloc(1, imx - 1, 1, jmx - 1, 1, kmx, nv);
  xl.alloc(0, imx + 1, 1, jmx - 1, 1, kmx - 1, nv); xr = xl;
  yl.alloc(1, imx - 1, 0, jmx + 1, 1, kmx - 1, nv); yr = yl;
  zl.alloc(1, imx - 1, 1, jmx - 1, 0, kmx + 1, nv); zr = zl;
  delta_t.alloc(1, imx - 1, 1, jmx - 1, 1, kmx - 1);
  pdif.alloc(0, imx, 0, jmx, 0, kmx);
  cells.alloc(imx + 2, jmx + 2, kmx + 2);
  If.alloc(imx + 3, jmx + 2, kmx + 2);
  Jf.alloc(imx + 2, jmx + 3, kmx + 2);
  Kf.alloc(imx + 2, jmx + 2, kmx + 3);
  if (c.mu_ref != 0.0) {
    gx.alloc(0, imx, 0, jmx, 0, kmx, n_grad); gy = gx; gz = gx;
    mu.alloc(-2, imx + 2, -2, jmx + 2, -2, kmx + 2, c.mu_ref);  // viscosity.f90:527 "mu = flow%mu_ref"
  }
  if (c.turbulence != ORC_TURB_NONE) {
    mu_t.alloc(-2, imx + 2, -2, jmx + 2, -2, kmx + 2);
    F1.alloc(-2, imx + 2, -2, jmx + 2, -2, kmx + 2);
    dist.alloc(-2, imx + 2, -2, jmx + 2, -2, kmx + 2);
  }
  if (c.transition == 2) dvdy.alloc(-2, imx + 2, -2, jmx + 2, -2, kmx + 2);   // CC.f90:59-66
  if (c.time_accuracy != ORC_T_NONE) U_store.alloc(-2, imx + 2, -2, jmx + 2, -2, kmx + 2, nv);
  if (c.time_accuracy == ORC_T_RK2 || c.time_accuracy == ORC_T_RK4)
    R_store.alloc(1, imx - 1, 1, jmx - 1, 1, kmx - 1, nv);
  // bc.f90:48-66
  c2 = 1 + c.accur; c3 = 0.5 * c.accur; c1 = c2 - c3;
  zF.assign(imx + 1, 1); zG.assign(jmx + 1, 1); zH.assign(kmx + 1, 1);
  auto wallish = [](int id) { return id == -5 || id == -6 || id == -7; };
  if (wallish(c.bc_id[0])) zF[1] = 0;
  if (wallish(c.bc_id[2])) zG[1] = 0;
  if (wallish(c.bc_id[4])) zH[1] = 0;
  if (wallish(c.bc_id[1])) zF[imx] = 0;
  if (wallish(c.bc_id[3])) zG[jmx] = 0;
  if (wallish(c.bc_id[5])) zH[kmx] = 0;
  // global_sst.f90:15-16
  gama1 = (beta1 / bstar) - ((sigma_w1 * (kappa_sst * kappa_sst)) / std::sqrt(bstar));
  gama2 = (beta2 / bstar) - ((sigma_w2 * (kappa_sst * kappa_sst)) / std::sqrt(bstar));
  ppm_flag = 0;
  error = 0;
}

// update.f90:170  Temp = qp(:,:,:,5)/(R_gas*qp(:,:,:,1)) over the whole array
void Block::refresh_temp() {
  for (int k = -2; k <= kmx + 2; ++k)
    for (int j = -2; j <= jmx + 2; ++j)
      for (int i = -2; i <= imx + 2; ++i) Temp(i, j, k) = qp(i, j, k, 5) / (c.R_gas * qp(i, j, k, 1));
}

// ---------------------------------------------------------------------------------------------
// Ghost-cell boundary conditions.  A "face frame" maps (layer, a, b) to (i,j,k):
//   face 1 imin: cell(l) interior layer l=1..3 -> i=l ; ghost layer l -> i=1-l ; transverse (a,b)=(j,k)
struct Frame {
  int face;  // 1..6
  int imx, jmx, kmx;
  int na, nb;  // interior transverse extents
  inline void interior(int l, int a, int b, int& i, int& j, int& k) const {
    switch (face) {
      case 1: i = l; j = a; k = b; break;
      case 2: i = imx - l; j = a; k = b; break;
      case 3: i = a; j = l; k = b; break;
      case 4: i = a; j = jmx - l; k = b; break;
      case 5: i = a; j = b; k = l; break;
      default: i = a; j = b; k = kmx - l; break;
    }
  }
  inline void ghost(int l, int a, int b, int& i, int& j, int& k) const {
    switch (face) {
      case 1: i = 1 - l; j = a; k = b; break;
      case 2: i = imx + l - 1; j = a; k = b; break;
      case 3: i = a; j = 1 - l; k = b; break;
      case 4: i = a; j = jmx + l - 1; k = b; break;
      case 5: i = a; j = b; k = 1 - l; break;
      default: i = a; j = b; k = kmx + l - 1; break;
    }
  }
};

static Frame make_frame(const Block& B, int face) {
  Frame f; f.face = face; f.imx = B.imx; f.jmx = B.jmx; f.kmx = B.kmx;
  if (face <= 2) { f.na = B.jmx - 1; f.nb = B.kmx - 1; }
  else if (face <= 4) { f.na = B.imx - 1; f.nb = B.kmx - 1; }
  else { f.na = B.imx - 1; f.nb = B.jmx - 1; }
  return f;
}

// copy_bc.f90:58-131  copy3(var, type, face)
enum { FLAT, SYMM, ANTI };
static void copy3(Block& B, int var, int type, int face) {
  Frame f = make_frame(B, face);
  double a1, a2, a3; int s[3];
  if (type == ANTI) { a1 = 1.; a2 = -1.; a3 = 0.; s[0] = 1; s[1] = 2; s[2] = 3; }
  else if (type == FLAT) { a1 = 1.; a2 = 1.; a3 = 0.; s[0] = 1; s[1] = 1; s[2] = 1; }
  else { a1 = B.c1; a2 = B.c2; a3 = B.c3; s[0] = 1; s[1] = 2; s[2] = 3; }
  for (int l = 1; l <= 3; ++l)
    for (int b = 1; b <= f.nb; ++b)
      for (int a = 1; a <= f.na; ++a) {
        int i, j, k, i1, j1, k1, i2, j2, k2;
        f.ghost(l, a, b, i, j, k);
        f.interior(s[l - 1], a, b, i1, j1, k1);
        f.interior(s[l - 1] + 1, a, b, i2, j2, k2);
        B.qp(i, j, k, var) = (a2 * B.qp(i1, j1, k1, var) - a3 * B.qp(i2, j2, k2, var)) / a1;
      }
}

// bc_primitive.f90:484-526  fix(var, fix_val, face)
static void fix(Block& B, int var, int slot, int face) {
  Frame f = make_frame(B, face);
  double v = B.c.fixed[slot][face - 1];
  for (int l = 1; l <= 3; ++l)
    for (int b = 1; b <= f.nb; ++b)
      for (int a = 1; a <= f.na; ++a) {
        int i, j, k; f.ghost(l, a, b, i, j, k);
        B.qp(i, j, k, var) = v;
      }
}

static inline bool is_sst(const Block& B) { return B.c.turbulence == ORC_TURB_SST || B.c.turbulence == ORC_TURB_SST2003; }
static inline bool is_sa(const Block& B) { return B.c.turbulence == ORC_TURB_SA; }
static inline bool is_kkl(const Block& B) { return B.c.turbulence == ORC_TURB_KKL; }
static inline bool is_lctm(const Block& B) { return B.c.transition == 2; }   // intermittency = variable 8 (bc_primitive.f90:143)

// FT_bc.f90:15-107  flow_tangency.  NOTE (kept defect): for J and K faces the dot product uses the
// Jfaces/Kfaces normal but the reflection subtracts 2*dot*Ifaces(i,1,k)%n etc. (FT_bc.f90:66-69,...)
static void flow_tangency(Block& B, int face) {
  Frame f = make_frame(B, face);
  for (int b = 1; b <= f.nb; ++b)
    for (int a = 1; a <= f.na; ++a)
      for (int l = 1; l <= 3; ++l) {
        int i, j, k, ig, jg, kg, fi, fj, fk;
        f.interior(l, a, b, i, j, k);
        f.ghost(l, a, b, ig, jg, kg);
        // boundary face index: layer 1 interior cell for min faces, ghost-side index for max faces
        switch (face) {
          case 1: fi = 1; fj = a; fk = b; break;
          case 2: fi = B.imx; fj = a; fk = b; break;
          case 3: fi = a; fj = 1; fk = b; break;
          case 4: fi = a; fj = B.jmx; fk = b; break;
          case 5: fi = a; fj = b; fk = 1; break;
          default: fi = a; fj = b; fk = B.kmx; break;
        }
        const Rec4& Fd = (face <= 2) ? B.If : (face <= 4 ? B.Jf : B.Kf);
        double dot = B.qp(i, j, k, 2) * Fd.nx(fi, fj, fk) + B.qp(i, j, k, 3) * Fd.ny(fi, fj, fk) +
                     B.qp(i, j, k, 4) * Fd.nz(fi, fj, fk);
        B.qp(ig, jg, kg, 2) = B.qp(i, j, k, 2) - (2.0 * dot * B.If.nx(fi, fj, fk));
        B.qp(ig, jg, kg, 3) = B.qp(i, j, k, 3) - (2.0 * dot * B.If.ny(fi, fj, fk));
        B.qp(ig, jg, kg, 4) = B.qp(i, j, k, 4) - (2.0 * dot * B.If.nz(fi, fj, fk));
      }
}

// bc_primitive.f90:564-643  set_omega_at_wall
static void set_omega_at_wall(Block& B, int face) {
  Frame f = make_frame(B, face);
  const OracleConfig& c = B.c;
  for (int l = 1; l <= 3; ++l)
    for (int b = 1; b <= f.nb; ++b)
      for (int a = 1; a <= f.na; ++a) {
        int i0, j0, k0, i1, j1, k1, il, jl, kl, ig, jg, kg;
        f.ghost(1, a, b, i0, j0, k0);      // first ghost
        f.interior(1, a, b, i1, j1, k1);   // first interior
        f.interior(l, a, b, il, jl, kl);
        f.ghost(l, a, b, ig, jg, kg);
        double T_face = 0.5 * ((B.qp(i0, j0, k0, 5) / B.qp(i0, j0, k0, 1)) + (B.qp(i1, j1, k1, 5) / B.qp(i1, j1, k1, 1))) / c.R_gas;
        double mu = c.mu_ref * std::pow(T_face / c.T_ref, 1.5) * ((c.T_ref + c.Sutherland_temp) / (T_face + c.Sutherland_temp));
        double rho = 0.5 * (B.qp(i0, j0, k0, 1) + B.qp(i1, j1, k1, 1));
        double d = 2 * B.dist(i1, j1, k1);
        B.qp(ig, jg, kg, 7) = 120 * mu / (rho * beta1 * (d * d)) - B.qp(il, jl, kl, 7);
      }
}

// bc_primitive.f90:1778-1946  temp_based_density
static void temp_based_density(Block& B, int face) {
  Frame f = make_frame(B, face);
  const OracleConfig& c = B.c;
  double T = c.fixed[ORC_FIX_WALL_TEMP][face - 1];
  if (T < 0.0) {
    for (int b = 1; b <= f.nb; ++b)
      for (int a = 1; a <= f.na; ++a) {
        int i, j, k; f.interior(1, a, b, i, j, k);
        double p = B.qp(i, j, k, 5), r = B.qp(i, j, k, 1);
        double stag_temp = (p / (c.R_gas * r)) * (1 + (0.5 * (c.gm - 1.) * c.gm * p / r));
        for (int l = 1; l <= 3; ++l) {
          int ig, jg, kg; f.ghost(l, a, b, ig, jg, kg);
          B.qp(ig, jg, kg, 1) = B.qp(ig, jg, kg, 5) / (c.R_gas * stag_temp);
        }
      }
  } else if (T > 1.0) {
    for (int b = 1; b <= f.nb; ++b)
      for (int a = 1; a <= f.na; ++a)
        for (int l = 1; l <= 3; ++l) {
          int i, j, k, ig, jg, kg;
          f.interior(l, a, b, i, j, k);
          f.ghost(l, a, b, ig, jg, kg);
          B.qp(ig, jg, kg, 1) = B.qp(ig, jg, kg, 5) / (c.R_gas * (2 * T - (B.qp(i, j, k, 5) / (c.R_gas * B.qp(i, j, k, 1)))));
        }
  } else {
    copy3(B, 1, SYMM, face);
  }
}

// bc_primitive.f90:645-1234  far_field (Riemann invariants).  The whole-face copy3/fix calls made
// from inside the per-cell loop (:700-757) are re-stated literally, including the flag logic.
static void ghost_plane_copy(Block& B, int face);
static void far_field(Block& B, int face) {
  Frame f = make_frame(B, face);
  const OracleConfig& c = B.c;
  const Rec4& Fd = (face <= 2) ? B.If : (face <= 4 ? B.Jf : B.Kf);
  const double sgn = (face % 2 == 1) ? -1.0 : 1.0;  // outward normal = -n on min faces
  int already_fixed = 0;
  for (int b = 1; b <= f.nb; ++b)
    for (int a = 1; a <= f.na; ++a) {
      int i, j, k, ig, jg, kg, fi, fj, fk;
      f.interior(1, a, b, i, j, k);
      f.ghost(1, a, b, ig, jg, kg);
      if (face % 2 == 1) { fi = i; fj = j; fk = k; } else { fi = ig; fj = jg; fk = kg; }
      double nx = sgn * Fd.nx(fi, fj, fk), ny = sgn * Fd.ny(fi, fj, fk), nz = sgn * Fd.nz(fi, fj, fk);
      double u = B.qp(i, j, k, 2), v = B.qp(i, j, k, 3), w = B.qp(i, j, k, 4);
      double uf = c.x_speed_inf, vf = c.y_speed_inf, wf = c.z_speed_inf;
      double cexp = std::sqrt(c.gm * B.qp(i, j, k, 5) / B.qp(i, j, k, 1));
      double cinf = std::sqrt(c.gm * c.pressure_inf / c.density_inf);
      double Unexp = u * nx + v * ny + w * nz;
      double Uninf = uf * nx + vf * ny + wf * nz;
      double Rinf = Uninf - 2 * cinf / (c.gm - 1.);
      double Rexp = Unexp + 2 * cexp / (c.gm - 1.);
      double Unb = 0.5 * (Rexp + Rinf);
      double Cb = 0.25 * (c.gm - 1.) * (Rexp - Rinf);
      if (Unb > 0.) {
        double vel_diff = Unb - Unexp;
        B.qp(ig, jg, kg, 2) = B.qp(i, j, k, 2) + vel_diff * nx;
        B.qp(ig, jg, kg, 3) = B.qp(i, j, k, 3) + vel_diff * ny;
        B.qp(ig, jg, kg, 4) = B.qp(i, j, k, 4) + vel_diff * nz;
        double s = B.qp(i, j, k, 5) / std::pow(B.qp(i, j, k, 1), c.gm);
        B.qp(ig, jg, kg, 1) = std::pow(Cb * Cb / (c.gm * s), 1. / (c.gm - 1.));
        B.qp(ig, jg, kg, 5) = (B.qp(ig, jg, kg, 1) * Cb * Cb / c.gm);
        if (is_sst(B) || is_kkl(B)) { copy3(B, 6, FLAT, face); copy3(B, 7, FLAT, face); }
        if (is_sa(B)) copy3(B, 6, FLAT, face);
        if (is_lctm(B)) copy3(B, 8, FLAT, face);
        already_fixed = 0;
      } else {
        double vel_diff = Unb - Uninf;
        B.qp(ig, jg, kg, 2) = c.x_speed_inf + vel_diff * nx;
        B.qp(ig, jg, kg, 3) = c.y_speed_inf + vel_diff * ny;
        B.qp(ig, jg, kg, 4) = c.z_speed_inf + vel_diff * nz;
        double s = c.pressure_inf / std::pow(c.density_inf, c.gm);
        B.qp(ig, jg, kg, 1) = std::pow(Cb * Cb / (c.gm * s), 1. / (c.gm - 1.));
        B.qp(ig, jg, kg, 5) = (B.qp(ig, jg, kg, 1) * Cb * Cb / c.gm);
        if (already_fixed == 0) {
          if (is_sst(B)) { fix(B, 6, ORC_FIX_TK, face); fix(B, 7, ORC_FIX_TW, face); }
          if (is_kkl(B)) { fix(B, 6, ORC_FIX_TK, face); fix(B, 7, ORC_FIX_TKL, face); }
          if (is_sa(B)) fix(B, 6, ORC_FIX_TV, face);
          if (is_lctm(B)) fix(B, 8, ORC_FIX_TGM, face);
        }
        already_fixed = 1;
      }
    }
  ghost_plane_copy(B, face);
}

// qp(-1,:,:,:) = qp(0,:,:,:) ; qp(-2,:,:,:) = qp(0,:,:,:)  -- whole planes incl. ghost rows (:762-763, :1344-1345)
static void ghost_plane_copy(Block& B, int face) {
  int lo[3] = {-2, -2, -2}, hi[3] = {B.imx + 2, B.jmx + 2, B.kmx + 2};
  int ax = (face - 1) / 2, t1 = (ax + 1) % 3, t2 = (ax + 2) % 3;
  int g0 = (face % 2 == 1) ? 0 : hi[ax] - 2;          // first ghost index along ax
  int step = (face % 2 == 1) ? -1 : 1;
  for (int l = 1; l <= B.nv; ++l)
    for (int m = 1; m <= 2; ++m) {
      int idx[3], src[3];
      for (int q = lo[t2]; q <= hi[t2]; ++q)
        for (int p = lo[t1]; p <= hi[t1]; ++p) {
          idx[ax] = g0 + m * step; idx[t1] = p; idx[t2] = q;
          src[ax] = g0; src[t1] = p; src[t2] = q;
          B.qp(idx[0], idx[1], idx[2], l) = B.qp(src[0], src[1], src[2], l);
        }
    }
}

// bc_primitive.f90:1237-1776  total_pressure (id -11): far-field Riemann velocity, then p from the fixed total
// pressure at the boundary Mach number and rho = gm*p/Cb^2.  Mb is taken from the ghost velocity on every face except
// kmin, which reads the INTERIOR cell (:1678 "x_speed(i,j,k)" instead of "(i,j,k-1)") -- reproduced.  The whole-face
// copy3 / fix calls sit inside the per-cell loop without the far-field's already_fixed flag (:1286-1336).
static void total_pressure(Block& B, int face) {
  Frame f = make_frame(B, face);
  const OracleConfig& c = B.c;
  const Rec4& Fd = (face <= 2) ? B.If : (face <= 4 ? B.Jf : B.Kf);
  const double sgn = (face % 2 == 1) ? -1.0 : 1.0;  // outward normal = -n on min faces
  for (int b = 1; b <= f.nb; ++b)
    for (int a = 1; a <= f.na; ++a) {
      int i, j, k, ig, jg, kg, fi, fj, fk;
      f.interior(1, a, b, i, j, k);
      f.ghost(1, a, b, ig, jg, kg);
      if (face % 2 == 1) { fi = i; fj = j; fk = k; } else { fi = ig; fj = jg; fk = kg; }
      double nx = sgn * Fd.nx(fi, fj, fk), ny = sgn * Fd.ny(fi, fj, fk), nz = sgn * Fd.nz(fi, fj, fk);
      double u = B.qp(i, j, k, 2), v = B.qp(i, j, k, 3), w = B.qp(i, j, k, 4);
      double uf = c.x_speed_inf, vf = c.y_speed_inf, wf = c.z_speed_inf;
      double cexp = std::sqrt(c.gm * B.qp(i, j, k, 5) / B.qp(i, j, k, 1));
      double cinf = std::sqrt(c.gm * c.pressure_inf / c.density_inf);
      double Unexp = u * nx + v * ny + w * nz;
      double Uninf = uf * nx + vf * ny + wf * nz;
      double Rinf = Uninf - 2 * cinf / (c.gm - 1.);
      double Rexp = Unexp + 2 * cexp / (c.gm - 1.);
      double Unb = 0.5 * (Rexp + Rinf);
      double Cb = 0.25 * (c.gm - 1.) * (Rexp - Rinf);
      if (Unb > 0.) {
        double vel_diff = Unb - Unexp;
        B.qp(ig, jg, kg, 2) = B.qp(i, j, k, 2) + vel_diff * nx;
        B.qp(ig, jg, kg, 3) = B.qp(i, j, k, 3) + vel_diff * ny;
        B.qp(ig, jg, kg, 4) = B.qp(i, j, k, 4) + vel_diff * nz;
        if (is_sst(B) || is_kkl(B)) { copy3(B, 6, FLAT, face); copy3(B, 7, FLAT, face); }
        if (is_sa(B)) copy3(B, 6, FLAT, face);
        if (is_lctm(B)) copy3(B, 8, FLAT, face);
      } else {
        double vel_diff = Unb - Uninf;
        B.qp(ig, jg, kg, 2) = c.x_speed_inf + vel_diff * nx;
        B.qp(ig, jg, kg, 3) = c.y_speed_inf + vel_diff * ny;
        B.qp(ig, jg, kg, 4) = c.z_speed_inf + vel_diff * nz;
        if (is_sst(B)) { fix(B, 6, ORC_FIX_TK, face); fix(B, 7, ORC_FIX_TW, face); }
        if (is_kkl(B)) { fix(B, 6, ORC_FIX_TK, face); fix(B, 7, ORC_FIX_TKL, face); }
        if (is_sa(B)) fix(B, 6, ORC_FIX_TV, face);
        if (is_lctm(B)) fix(B, 8, ORC_FIX_TGM, face);
      }
      int im = ig, jm = jg, km = kg;
      if (face == 5) { im = i; jm = j; km = k; }
      double Mb = std::sqrt(B.qp(im, jm, km, 2) * B.qp(im, jm, km, 2) + B.qp(im, jm, km, 3) * B.qp(im, jm, km, 3) +
                            B.qp(im, jm, km, 4) * B.qp(im, jm, km, 4)) / Cb;
      B.qp(ig, jg, kg, 5) = c.fixed[ORC_FIX_TPRESSURE][face - 1] / std::pow((1 + 0.5 * (c.gm - 1.) * Mb * Mb), c.gm / (c.gm - 1.));
      B.qp(ig, jg, kg, 1) = c.gm * B.qp(ig, jg, kg, 5) / (Cb * Cb);
    }
  ghost_plane_copy(B, face);
}

// bc_primitive.f90:1948-1978 periodic_bc (single-block periodicity, id -9): whole slabs
static void periodic_bc(Block& B, int face) {
  int lo[3] = {-2, -2, -2}, hi[3] = {B.imx + 2, B.jmx + 2, B.kmx + 2};
  int mx[3] = {B.imx, B.jmx, B.kmx};
  int ax = (face - 1) / 2;
  for (int l = 1; l <= B.nv; ++l) {
    int idx[3];
    for (idx[2] = lo[2]; idx[2] <= hi[2]; ++idx[2])
      for (idx[1] = lo[1]; idx[1] <= hi[1]; ++idx[1])
        for (idx[0] = lo[0]; idx[0] <= hi[0]; ++idx[0]) {
          int src[3] = {idx[0], idx[1], idx[2]};
          if (face % 2 == 1) {  // qp(-2:0) = qp(mx-3:mx-1)
            if (idx[ax] > 0) continue;
            src[ax] = idx[ax] + mx[ax] - 1;
          } else {              // qp(mx:mx+2) = qp(1:3)
            if (idx[ax] < mx[ax]) continue;
            src[ax] = idx[ax] - mx[ax] + 1;
          }
          B.qp(idx[0], idx[1], idx[2], l) = B.qp(src[0], src[1], src[2], l);
        }
  }
}

// bc_primitive.f90:55-226 populate_ghost_primitive
void Block::populate_ghost_primitive() {
  Block& B = *this;
  const bool sst = is_sst(B);
  const bool sa = is_sa(B);   // the SA variable follows the pattern of tk on every face (bc_primitive.f90:246,288,327,373,425,463,543)
  // k-kL: kL (variable 7) follows the pattern of omega, except that the wall takes the anti copy instead of the wall-omega rule (:548-550)
  // and the subsonic inlet fixes it to fixed_tw, not fixed_tkl (:333-336, reproduced)
  const bool kkl = is_kkl(B);
  // lctm2015: the intermittency is fixed at the two inlets (:259-265, :340-346) and copied flat everywhere else (:299-304 ... :554-559)
  const bool lctm = is_lctm(B);
  for (int face = 1; face <= 6; ++face) {
    switch (c.bc_id[face - 1]) {
      case -1:  // supersonic_inlet :229
        if (current_iter <= 2) {
          fix(B, 1, ORC_FIX_DENSITY, face); fix(B, 2, ORC_FIX_X_SPEED, face); fix(B, 3, ORC_FIX_Y_SPEED, face);
          fix(B, 4, ORC_FIX_Z_SPEED, face); fix(B, 5, ORC_FIX_PRESSURE, face);
          if (sst) { fix(B, 6, ORC_FIX_TK, face); fix(B, 7, ORC_FIX_TW, face); }
          if (kkl) { fix(B, 6, ORC_FIX_TK, face); fix(B, 7, ORC_FIX_TKL, face); }
          if (sa) fix(B, 6, ORC_FIX_TV, face);
          if (lctm) fix(B, 8, ORC_FIX_TGM, face);
        }
        break;
      case -2:  // supersonic_outlet :270
        for (int v = 1; v <= 5; ++v) copy3(B, v, FLAT, face);
        if (sst || kkl) { copy3(B, 6, FLAT, face); copy3(B, 7, FLAT, face); }
        if (sa) copy3(B, 6, FLAT, face);
        if (lctm) copy3(B, 8, FLAT, face);
        break;
      case -3:  // subsonic_inlet :308
        if (current_iter <= 2) {
          fix(B, 1, ORC_FIX_DENSITY, face); fix(B, 2, ORC_FIX_X_SPEED, face); fix(B, 3, ORC_FIX_Y_SPEED, face);
          fix(B, 4, ORC_FIX_Z_SPEED, face);
          if (sst || kkl) { fix(B, 6, ORC_FIX_TK, face); fix(B, 7, ORC_FIX_TW, face); }
          if (sa) fix(B, 6, ORC_FIX_TV, face);
          if (lctm) fix(B, 8, ORC_FIX_TGM, face);
        }
        copy3(B, 5, FLAT, face);
        break;
      case -4:  // subsonic_outlet :352
        for (int v = 1; v <= 4; ++v) copy3(B, v, FLAT, face);
        if (current_iter <= 2) fix(B, 5, ORC_FIX_PRESSURE, face);
        if (sst || kkl) { copy3(B, 6, FLAT, face); copy3(B, 7, FLAT, face); }
        if (sa) copy3(B, 6, FLAT, face);
        if (lctm) copy3(B, 8, FLAT, face);
        break;
      case -5:  // wall :392 -> pressure symm, temp_based_density, no_slip :528
        copy3(B, 5, SYMM, face);
        temp_based_density(B, face);
        copy3(B, 2, ANTI, face); copy3(B, 3, ANTI, face); copy3(B, 4, ANTI, face);
        if (sst) { copy3(B, 6, ANTI, face); set_omega_at_wall(B, face); }
        if (kkl) { copy3(B, 6, ANTI, face); copy3(B, 7, ANTI, face); }
        if (sa) copy3(B, 6, ANTI, face);
        if (lctm) copy3(B, 8, FLAT, face);
        break;
      case -6:  // slip_wall :405
        copy3(B, 1, SYMM, face); copy3(B, 5, SYMM, face);
        if (sst || kkl) { copy3(B, 6, SYMM, face); copy3(B, 7, SYMM, face); }
        if (sa) copy3(B, 6, SYMM, face);
        if (lctm) copy3(B, 8, FLAT, face);
        flow_tangency(B, face);
        break;
      case -7:  // pole :446
        for (int v = 1; v <= 5; ++v) copy3(B, v, FLAT, face);
        if (sst || kkl) { copy3(B, 6, FLAT, face); copy3(B, 7, FLAT, face); }
        if (sa) copy3(B, 6, FLAT, face);
        if (lctm) copy3(B, 8, FLAT, face);
        break;
      case -8: far_field(B, face); break;
      case -9: periodic_bc(B, face); break;
      case -11: total_pressure(B, face); break;
      default: break;  // interface (>=0) or -10
    }
  }
  // edge / corner fill, exact statement order of bc_primitive.f90:209-224 (factor 0.33, whole slices)
  for (int l = 1; l <= nv; ++l) {
    for (int i = -2; i <= imx + 2; ++i) qp(i, 0, 0, l) = 0.33 * (qp(i, 1, 1, l) + qp(i, 0, 1, l) + qp(i, 1, 0, l));
  }
  for (int l = 1; l <= nv; ++l)
    for (int i = -2; i <= imx + 2; ++i) qp(i, 0, kmx, l) = 0.33 * (qp(i, 1, kmx - 1, l) + qp(i, 0, kmx - 1, l) + qp(i, 1, kmx, l));
  for (int l = 1; l <= nv; ++l)
    for (int i = -2; i <= imx + 2; ++i) qp(i, jmx, 0, l) = 0.33 * (qp(i, jmx - 1, 1, l) + qp(i, jmx, 1, l) + qp(i, jmx - 1, 0, l));
  for (int l = 1; l <= nv; ++l)
    for (int i = -2; i <= imx + 2; ++i) qp(i, jmx, kmx, l) = 0.33 * (qp(i, jmx - 1, kmx - 1, l) + qp(i, jmx, kmx - 1, l) + qp(i, jmx - 1, kmx, l));
  for (int l = 1; l <= nv; ++l)
    for (int k = -2; k <= kmx + 2; ++k) qp(imx, 0, k, l) = 0.33 * (qp(imx - 1, 1, k, l) + qp(imx - 1, 0, k, l) + qp(imx, 1, k, l));
  for (int l = 1; l <= nv; ++l)
    for (int k = -2; k <= kmx + 2; ++k) qp(0, 0, k, l) = 0.33 * (qp(1, 1, k, l) + qp(1, 0, k, l) + qp(0, 1, k, l));
  for (int l = 1; l <= nv; ++l)
    for (int k = -2; k <= kmx + 2; ++k) qp(0, jmx, k, l) = 0.33 * (qp(1, jmx - 1, k, l) + qp(1, jmx, k, l) + qp(0, jmx - 1, k, l));
  for (int l = 1; l <= nv; ++l)
    for (int k = -2; k <= kmx + 2; ++k) qp(imx, jmx, k, l) = 0.33 * (qp(imx - 1, jmx - 1, k, l) + qp(imx - 1, jmx, k, l) + qp(imx, jmx - 1, k, l));
  for (int l = 1; l <= nv; ++l) qp(0, 0, 0, l) = 0.33 * (qp(1, 0, 0, l) + qp(0, 1, 0, l) + qp(0, 0, 1, l));
  for (int l = 1; l <= nv; ++l) qp(imx, 0, 0, l) = 0.33 * (qp(imx - 1, 0, 0, l) + qp(imx, 1, 0, l) + qp(imx, 0, 1, l));
  for (int l = 1; l <= nv; ++l) qp(0, jmx, 0, l) = 0.33 * (qp(1, jmx, 0, l) + qp(0, jmx - 1, 0, l) + qp(0, jmx, 1, l));
  for (int l = 1; l <= nv; ++l) qp(0, 0, kmx, l) = 0.33 * (qp(1, 0, kmx, l) + qp(0, 1, kmx, l) + qp(0, 0, kmx - 1, l));
  for (int l = 1; l <= nv; ++l) qp(imx, jmx, 0, l) = 0.33 * (qp(imx - 1, jmx, 0, l) + qp(imx, jmx - 1, 0, l) + qp(imx, jmx, 1, l));
  for (int l = 1; l <= nv; ++l) qp(imx, 0, kmx, l) = 0.33 * (qp(imx - 1, 0, kmx, l) + qp(imx, 1, kmx, l) + qp(imx, 0, kmx - 1, l));
  for (int l = 1; l <= nv; ++l) qp(0, jmx, kmx, l) = 0.33 * (qp(1, jmx, kmx, l) + qp(0, jmx - 1, kmx, l) + qp(0, jmx, kmx - 1, l));
  for (int l = 1; l <= nv; ++l) qp(imx, jmx, kmx, l) = 0.33 * (qp(imx - 1, jmx, kmx, l) + qp(imx, jmx - 1, kmx, l) + qp(imx, jmx, kmx - 1, l));
}

// ---------------------------------------------------------------------------------------------
// Face reconstruction.  All variants loop cells 1-ii .. imx-1+ii in the sweep direction and write
// left(i+ii) / right(i)  (muscl.f90:161-196, weno.f90:55-91, weno_NM.f90:66-118, ppm.f90:44-105).

// muscl.f90:115-200 compute_face_state
static void muscl_dir(const Block& B, Arr4& fl, Arr4& fr, int ii, int jj, int kk, int lam_switch, int turb_switch) {
  const double alpha = 2. / 3., phi = 1.0, kappa = 1. / 3., eps = 1e-14;
  int switch_L = lam_switch;
  for (int l = 1; l <= B.nv; ++l) {
    if (l >= 6) switch_L = turb_switch;
    for (int k = 1 - kk; k <= B.kmx - 1 + kk; ++k)
      for (int j = 1 - jj; j <= B.jmx - 1 + jj; ++j)
        for (int i = 1 - ii; i <= B.imx - 1 + ii; ++i) {
          double fd = B.qp(i + ii, j + jj, k + kk, l) - B.qp(i, j, k, l);
          double bd = B.qp(i, j, k, l) - B.qp(i - ii, j - jj, k - kk, l);
          double r = fd / (bd + fsign(eps, bd));
          double psi1 = std::fmax(0., fmin3(2 * r, alpha * (r - 1.0) + 1.0, 2.));
          r = bd / (fd + fsign(eps, fd));
          double psi2 = std::fmax(0., fmin3(2 * r, alpha * (r - 1.0) + 1.0, 2.));
          psi1 = (1 - (1 - psi1) * switch_L);
          psi2 = (1 - (1 - psi2) * switch_L);
          fl(i + ii, j + jj, k + kk, l) = B.qp(i, j, k, l) + 0.25 * phi * (((1. - kappa) * psi1 * bd) + ((1. + kappa) * psi2 * fd));
          fr(i, j, k, l) = B.qp(i, j, k, l) - 0.25 * phi * (((1. + kappa) * psi1 * bd) + ((1. - kappa) * psi2 * fd));
        }
  }
}

// muscl.f90:37-112 / ppm.f90:108-170 pressure_based_switching: pdif of every interior cell from its two neighbours along the
// direction, the two ghost positions copy the first / last interior value, then both face states are pulled towards the cell
// value.  pdif is a module array (0:imx,0:jmx,0:kmx) that persists between calls; everything read here is written here.
static void pressure_based_switching(Block& B, Arr4& fl, Arr4& fr, int ii, int jj, int kk) {
  Arr3& pdif = B.pdif;
  const int imx = B.imx, jmx = B.jmx, kmx = B.kmx;
  for (int k = 1; k <= kmx - 1; ++k)
    for (int j = 1; j <= jmx - 1; ++j)
      for (int i = 1; i <= imx - 1; ++i) {
        double pd2 = std::fabs(B.qp(i + ii, j + jj, k + kk, 5) - B.qp(i - ii, j - jj, k - kk, 5));
        pdif(i, j, k) = 1 - (pd2 / (pd2 + B.c.pressure_inf));
      }
  // ghost cells: the plane at index 0 along the direction takes the plane at 1, the plane at mx takes mx-1
  const int mx = ii ? imx : (jj ? jmx : kmx);
  for (int k = 1; k <= (kk ? 1 : kmx - 1); ++k)
    for (int j = 1; j <= (jj ? 1 : jmx - 1); ++j)
      for (int i = 1; i <= (ii ? 1 : imx - 1); ++i) {
        pdif(i - ii, j - jj, k - kk) = pdif(i, j, k);
        const int i2 = ii ? mx - 1 : i, j2 = jj ? mx - 1 : j, k2 = kk ? mx - 1 : k;
        pdif(i2 + ii, j2 + jj, k2 + kk) = pdif(i2, j2, k2);
      }
  for (int k = 1; k <= kmx - (1 - kk); ++k)
    for (int j = 1; j <= jmx - (1 - jj); ++j)
      for (int i = 1; i <= imx - (1 - ii); ++i)
        for (int l = 1; l <= B.nv; ++l) {
          fl(i, j, k, l) = B.qp(i - ii, j - jj, k - kk, l) + (pdif(i - ii, j - jj, k - kk) * (fl(i, j, k, l) - B.qp(i - ii, j - jj, k - kk, l)));
          fr(i, j, k, l) = B.qp(i, j, k, l) - (pdif(i, j, k) * (B.qp(i, j, k, l) - fr(i, j, k, l)));
        }
}

// weno.f90:24-93
static void weno_dir(const Block& B, Arr4& fl, Arr4& fr, int ii, int jj, int kk) {
  const double eps = 1e-6;
  const double g[3] = {1.0 / 10.0, 6.0 / 10.0, 3.0 / 10.0};
  for (int l = 1; l <= B.nv; ++l)
    for (int k = 1 - kk; k <= B.kmx - 1 + kk; ++k)
      for (int j = 1 - jj; j <= B.jmx - 1 + jj; ++j)
        for (int i = 1 - ii; i <= B.imx - 1 + ii; ++i) {
          double um2 = B.qp(i - 2 * ii, j - 2 * jj, k - 2 * kk, l), um1 = B.qp(i - ii, j - jj, k - kk, l);
          double u0 = B.qp(i, j, k, l), up1 = B.qp(i + ii, j + jj, k + kk, l), up2 = B.qp(i + 2 * ii, j + 2 * jj, k + 2 * kk, l);
          double P[3], Bt[3], w[3];
          P[0] = (2.0 * um2 - 7.0 * um1 + 11.0 * u0) / 6.0;
          P[1] = (-1.0 * um1 + 5.0 * u0 + 2.0 * up1) / 6.0;
          P[2] = (2.0 * u0 + 5.0 * up1 - 1.0 * up2) / 6.0;
          double t;
          t = (um2 - 2.0 * um1 + u0); double s = (um2 - 4.0 * um1 + 3.0 * u0);
          Bt[0] = (13.0 / 12.0) * (t * t) + (1.0 / 4.0) * (s * s);
          t = (um1 - 2.0 * u0 + up1); s = (um1 - up1);
          Bt[1] = (13.0 / 12.0) * (t * t) + (1.0 / 4.0) * (s * s);
          t = (u0 - 2.0 * up1 + up2); s = (3.0 * u0 - 4.0 * up1 + up2);
          Bt[2] = (13.0 / 12.0) * (t * t) + (1.0 / 4.0) * (s * s);
          for (int m = 0; m < 3; ++m) { double e = eps + Bt[m]; w[m] = g[m] / (e * e); }
          fl(i + ii, j + jj, k + kk, l) = ((w[0] * P[0] + w[1] * P[1]) + w[2] * P[2]) / ((w[0] + w[1]) + w[2]);
          P[0] = (2.0 * up2 - 7.0 * up1 + 11.0 * u0) / 6.0;
          P[1] = (-1.0 * up1 + 5.0 * u0 + 2.0 * um1) / 6.0;
          P[2] = (2.0 * u0 + 5.0 * um1 - 1.0 * um2) / 6.0;
          { double e = eps + Bt[2]; w[0] = g[0] / (e * e); }
          { double e = eps + Bt[1]; w[1] = g[1] / (e * e); }
          { double e = eps + Bt[0]; w[2] = g[2] / (e * e); }
          fr(i, j, k, l) = ((w[0] * P[0] + w[1] * P[1]) + w[2] * P[2]) / ((w[0] + w[1]) + w[2]);
        }
}

// weno_NM.f90:21-120 (volume-weighted non-uniform-mesh WENO)
static void weno_nm_dir(const Block& B, Arr4& fl, Arr4& fr, int ii, int jj, int kk) {
  const double eps = 1e-6;
  const double g[3] = {1.0 / 10.0, 6.0 / 10.0, 3.0 / 10.0};
  for (int l = 1; l <= B.nv; ++l)
    for (int k = 1 - kk; k <= B.kmx - 1 + kk; ++k)
      for (int j = 1 - jj; j <= B.jmx - 1 + jj; ++j)
        for (int i = 1 - ii; i <= B.imx - 1 + ii; ++i) {
          double U[5], V[5];
          for (int m = -2; m <= 2; ++m) {
            U[m + 2] = B.qp(i + m * ii, j + m * jj, k + m * kk, l);
            V[m + 2] = B.cells.vol(i + m * ii, j + m * jj, k + m * kk);
          }
          const double um2 = U[0], um1 = U[1], u0 = U[2], up1 = U[3], up2 = U[4];
          double alpha12 = V[4] / (V[3] + V[4]);
          double alpha01 = V[3] / (V[2] + V[3]);
          double alpha10 = V[2] / (V[1] + V[2]);
          double alpha21 = V[1] / (V[0] + V[1]);
          double U01 = (1.0 - alpha01) * u0 + alpha01 * up1;
          double U12 = (1.0 - alpha12) * up1 + alpha12 * up2;
          double U10 = (1.0 - alpha10) * um1 + alpha10 * u0;
          double U21 = (1.0 - alpha21) * um2 + alpha21 * um1;
          double U00 = um1 + (1.0 - alpha21) * (um1 - um2);
          double U11 = up1 + alpha12 * (up1 - up2);
          double P[3], Bt[3], w[3], t, s;
          P[0] = (6.0 * u0 - 1.0 * U10 - 2.0 * U00) / 3.0;
          P[1] = (-1.0 * U10 + 2.0 * u0 + 2.0 * U01) / 3.0;
          P[2] = (2.0 * U01 + 2.0 * up1 - 1.0 * U12) / 3.0;
          t = (2 * U10 - 2.0 * U00); s = (4 * u0 - 2.0 * U10 - 2.0 * U00);
          Bt[0] = (13.0 / 12.0) * (t * t) + (1.0 / 4.0) * (s * s);
          t = (2 * U10 - 4.0 * u0 + 2 * U01); s = (-2 * U10 + 2.0 * U01);
          Bt[1] = (13.0 / 12.0) * (t * t) + (1.0 / 4.0) * (s * s);
          t = (2 * U01 - 4.0 * up1 + 2 * U12); s = (-6 * U01 + 8.0 * up1 - 2.0 * U12);
          Bt[2] = (13.0 / 12.0) * (t * t) + (1.0 / 4.0) * (s * s);
          for (int m = 0; m < 3; ++m) { double e = eps + Bt[m]; w[m] = g[m] / (e * e); }
          fl(i + ii, j + jj, k + kk, l) = ((w[0] * P[0] + w[1] * P[1]) + w[2] * P[2]) / ((w[0] + w[1]) + w[2]);
          P[0] = (6.0 * u0 - 1.0 * U01 - 2.0 * U11) / 3.0;
          P[1] = (-1.0 * U01 + 2.0 * u0 + 2.0 * U10) / 3.0;
          P[2] = (2.0 * U10 + 2.0 * um1 - 1.0 * U21) / 3.0;
          t = (2 * U01 - 2.0 * U11); s = (4 * u0 - 2.0 * U01 - 2.0 * U11);
          Bt[0] = (13.0 / 12.0) * (t * t) + (1.0 / 4.0) * (s * s);
          t = (2 * U01 - 4.0 * u0 + 2 * U10); s = (-2 * U01 + 2.0 * U10);
          Bt[1] = (13.0 / 12.0) * (t * t) + (1.0 / 4.0) * (s * s);
          t = (2 * U10 - 4.0 * um1 + 2 * U21); s = (-6 * U10 + 8.0 * um1 - 2.0 * U21);
          Bt[2] = (13.0 / 12.0) * (t * t) + (1.0 / 4.0) * (s * s);
          for (int m = 0; m < 3; ++m) { double e = eps + Bt[m]; w[m] = g[m] / (e * e); }
          fr(i, j, k, l) = ((w[0] * P[0] + w[1] * P[1]) + w[2] * P[2]) / ((w[0] + w[1]) + w[2]);
        }
}

// ppm.f90:21-107: 4-point face estimate, then remove_extrema when the limiter switch is 1
static void ppm_dir(const Block& B, Arr4& fl, Arr4& fr, int ii, int jj, int kk, int limiter) {
  for (int l = 1; l <= B.nv; ++l)
    for (int k = 1 - kk; k <= B.kmx - 1 + 2 * kk; ++k)
      for (int j = 1 - jj; j <= B.jmx - 1 + 2 * jj; ++j)
        for (int i = 1 - ii; i <= B.imx - 1 + 2 * ii; ++i)
          fl(i, j, k, l) = (7. * (B.qp(i, j, k, l) + B.qp(i - ii, j - jj, k - kk, l)) -
                            (B.qp(i + ii, j + jj, k + kk, l) + B.qp(i - 2 * ii, j - 2 * jj, k - 2 * kk, l))) / 12.;
  fr.d = fl.d;
  if (limiter != 1) return;
  for (int l = 1; l <= B.nv; ++l)
    for (int k = 1 - kk; k <= B.kmx - 1 + kk; ++k)
      for (int j = 1 - jj; j <= B.jmx - 1 + jj; ++j)
        for (int i = 1 - ii; i <= B.imx - 1 + ii; ++i) {
          double& L = fl(i + ii, j + jj, k + kk, l);
          double& R = fr(i, j, k, l);
          double q = B.qp(i, j, k, l);
          if ((L - q) * (q - R) <= 0) { L = q; R = q; }
          else {
            double dqrl = L - R;
            double dq6 = 6. * (q - 0.5 * (L + R));
            if (dqrl * dq6 > dqrl * dqrl) R = 3. * q - 2. * L;
            else if (-dqrl * dqrl > dqrl * dq6) L = 3. * q - 2. * R;
          }
        }
}

// face_interpolant.f90:79-105
void Block::compute_face_interpolant() {
  switch (c.interpolant) {
    case ORC_NONE:  // extrapolate_cell_averages_to_faces :61-77
      for (int l = 1; l <= nv; ++l) {
        for (int k = 1; k <= kmx - 1; ++k) for (int j = 1; j <= jmx - 1; ++j) for (int i = 0; i <= imx + 1; ++i) { xl(i, j, k, l) = qp(i - 1, j, k, l); xr(i, j, k, l) = qp(i, j, k, l); }
        for (int k = 1; k <= kmx - 1; ++k) for (int j = 0; j <= jmx + 1; ++j) for (int i = 1; i <= imx - 1; ++i) { yl(i, j, k, l) = qp(i, j - 1, k, l); yr(i, j, k, l) = qp(i, j, k, l); }
        for (int k = 0; k <= kmx + 1; ++k) for (int j = 1; j <= jmx - 1; ++j) for (int i = 1; i <= imx - 1; ++i) { zl(i, j, k, l) = qp(i, j, k - 1, l); zr(i, j, k, l) = qp(i, j, k, l); }
      }
      break;
    case ORC_MUSCL:
      muscl_dir(*this, xl, xr, 1, 0, 0, c.limiter[0], c.tlimiter[0]);
      if (c.pb_switch[0] == 1) pressure_based_switching(*this, xl, xr, 1, 0, 0);   // muscl.f90:231-243
      muscl_dir(*this, yl, yr, 0, 1, 0, c.limiter[1], c.tlimiter[1]);
      if (c.pb_switch[1] == 1) pressure_based_switching(*this, yl, yr, 0, 1, 0);
      muscl_dir(*this, zl, zr, 0, 0, 1, c.limiter[2], c.tlimiter[2]);
      if (c.pb_switch[2] == 1) pressure_based_switching(*this, zl, zr, 0, 0, 1);
      break;
    case ORC_PPM:
      ppm_dir(*this, xl, xr, 1, 0, 0, c.limiter[0]);
      if (c.pb_switch[0] == 1) pressure_based_switching(*this, xl, xr, 1, 0, 0);   // ppm.f90:220-243
      ppm_dir(*this, yl, yr, 0, 1, 0, c.limiter[1]);
      if (c.pb_switch[1] == 1) pressure_based_switching(*this, yl, yr, 0, 1, 0);
      ppm_dir(*this, zl, zr, 0, 0, 1, c.limiter[2]);
      if (c.pb_switch[2] == 1) pressure_based_switching(*this, zl, zr, 0, 0, 1);
      break;
    case ORC_WENO:
      weno_dir(*this, xl, xr, 1, 0, 0); weno_dir(*this, yl, yr, 0, 1, 0); weno_dir(*this, zl, zr, 0, 0, 1);
      break;
    case ORC_WENO_NM:
      weno_nm_dir(*this, xl, xr, 1, 0, 0); weno_nm_dir(*this, yl, yr, 0, 1, 0); weno_nm_dir(*this, zl, zr, 0, 0, 1);
      break;
    default: error = 90; break;
  }
}

// boundary_state_reconstruction.f90:28-412.  Unguarded fd/bd (:104-108): Fortran min/max are re-stated
// with fmin/fmax so the 0/0 case (uniform field) selects the non-NaN branch.
static void recon_face(Block& B, int face) {
  Frame f = make_frame(B, face);
  const int ax = (face - 1) / 2;
  const int ii = ax == 0, jj = ax == 1, kk = ax == 2;
  Arr4& fl = ax == 0 ? B.xl : (ax == 1 ? B.yl : B.zl);
  Arr4& fr = ax == 0 ? B.xr : (ax == 1 ? B.yr : B.zr);
  const double phi = 1.0, kappa = 1. / 3.;
  const int id = B.c.bc_id[face - 1];
  if (B.ppm_flag == 1) {
    int switch_L = B.c.limiter[ax];
    for (int l = 1; l <= B.nv; ++l) {
      if (l >= 6) switch_L = B.c.tlimiter[ax];
      for (int b = 1; b <= f.nb; ++b)
        for (int a = 1; a <= f.na; ++a) {
          int i, j, k; f.interior(1, a, b, i, j, k);
          double fd = B.qp(i + ii, j + jj, k + kk, l) - B.qp(i, j, k, l);
          double bd = B.qp(i, j, k, l) - B.qp(i - ii, j - jj, k - kk, l);
          double r = fd / bd;
          double psi1 = std::fmax(0., fmin3(2 * r, (2 + r) / 3., 2.));
          psi1 = (1 - (1 - psi1) * switch_L);
          r = bd / fd;
          double psi2 = std::fmax(0., fmin3(2 * r, (2 + r) / 3., 2.));
          psi2 = (1 - (1 - psi2) * switch_L);
          fl(i + ii, j + jj, k + kk, l) = B.qp(i, j, k, l) + 0.25 * phi * (((1. - kappa) * psi1 * bd) + ((1. + kappa) * psi2 * fd));
          fr(i, j, k, l) = B.qp(i, j, k, l) - 0.25 * phi * (((1. + kappa) * psi1 * bd) + ((1. - kappa) * psi2 * fd));
        }
    }
  }
  for (int l = 1; l <= B.nv; ++l)
    for (int b = 1; b <= f.nb; ++b)
      for (int a = 1; a <= f.na; ++a) {
        int i, j, k, ig, jg, kg;
        f.interior(1, a, b, i, j, k);
        f.ghost(1, a, b, ig, jg, kg);
        // boundary-face index in the face-state arrays: min faces -> index of first interior cell,
        // max faces -> index of the first ghost cell
        int fi = (face % 2 == 1) ? i : ig, fj = (face % 2 == 1) ? j : jg, fk = (face % 2 == 1) ? k : kg;
        if (id == -8 || id == -9) {
          fl(fi, fj, fk, l) = B.qp(ig, jg, kg, l);
          fr(fi, fj, fk, l) = B.qp(ig, jg, kg, l);
        } else if (face % 2 == 1) {
          fl(fi, fj, fk, l) = 0.5 * (B.qp(ig, jg, kg, l) + B.qp(i, j, k, l));
        } else {
          fr(fi, fj, fk, l) = 0.5 * (B.qp(i, j, k, l) + B.qp(ig, jg, kg, l));
        }
      }
}

void Block::reconstruct_boundary_state() {
  if (c.interpolant == ORC_PPM || c.interpolant == ORC_WENO || c.interpolant == ORC_WENO_NM) ppm_flag = 1;
  for (int f = 0; f < 6; ++f) if (c.bc_id[f] == -7) ppm_flag = 1;
  if (c.interpolant != ORC_NONE)
    for (int face = 1; face <= 6; ++face)
      if (c.bc_id[face - 1] < 0 && c.bc_id[face - 1] != -10) recon_face(*this, face);
}

// ---------------------------------------------------------------------------------------------
// Inviscid fluxes: one face.  L,R = left/right primitive state (1..nv in [0..nv-1]).
// van_leer.f90:56-146, ldfss0.f90:57-161, ausm.f90:56-154, ausmP.f90:77-198, ausmUP.f90:88-216, slau.f90:82-204
static inline int supersonic_switch(double M) {  // max(0, 1 - floor(abs(M)))
  return (std::floor(std::fabs(M)) >= 1.0) ? 0 : 1;
}

void flux_kernel(int scheme, int nv, double gm, double MInf, const double* L, const double* R,
                 double A, double nx, double ny, double nz, int mask, double* out) {
  double Fp[8], Fm[8];
  if (scheme == ORC_VAN_LEER || scheme == ORC_LDFSS0 || scheme == ORC_AUSM) {
    double sound_speed_avg = 0.5 * (std::sqrt(gm * L[4] / L[0]) + std::sqrt(gm * R[4] / R[0]));
    double face_normal_speeds = L[1] * nx + L[2] * ny + L[3] * nz;
    double M_perp_left = face_normal_speeds / sound_speed_avg;
    double alpha_plus = 0.5 * (1.0 + fsign(1.0, M_perp_left));
    double beta_left = -(double)supersonic_switch(M_perp_left);
    double M_plus = 0.25 * ((1. + M_perp_left) * (1. + M_perp_left));
    double D_plus = 0.25 * ((1. + M_perp_left) * (1. + M_perp_left)) * (2. - M_perp_left);
    double c_plus = (alpha_plus * (1.0 + beta_left) * M_perp_left) - beta_left * M_plus;
    double scrD_plus = (alpha_plus * (1. + beta_left)) - (beta_left * D_plus);
    face_normal_speeds = R[1] * nx + R[2] * ny + R[3] * nz;
    double M_perp_right = face_normal_speeds / sound_speed_avg;
    double alpha_minus = 0.5 * (1.0 - fsign(1.0, M_perp_right));
    double beta_right = -(double)supersonic_switch(M_perp_right);
    double M_minus = -0.25 * ((1. - M_perp_right) * (1. - M_perp_right));
    double D_minus = 0.25 * ((1. - M_perp_right) * (1. - M_perp_right)) * (2. + M_perp_right);
    double c_minus = (alpha_minus * (1.0 + beta_right) * M_perp_right) - beta_right * M_minus;
    double scrD_minus = (alpha_minus * (1. + beta_right)) - (beta_right * D_minus);
    if (scheme == ORC_AUSM) {
      double temp_c = c_plus + c_minus;
      c_plus = std::fmax(0., temp_c);
      c_minus = std::fmin(0., temp_c);
    } else if (scheme == ORC_LDFSS0) {
      double t = (std::sqrt((M_perp_left * M_perp_left + M_perp_right * M_perp_right) * 0.5) - 1);
      double M_ldfss = 0.25 * beta_left * beta_right * (t * t);
      double M_plus_ldfss = M_ldfss * (1 - (L[4] - R[4]) / (2 * L[0] * (sound_speed_avg * sound_speed_avg)));
      double M_minus_ldfss = M_ldfss * (1 - (L[4] - R[4]) / (2 * R[0] * (sound_speed_avg * sound_speed_avg)));
      c_plus = c_plus - M_plus_ldfss;
      c_minus = c_minus + M_minus_ldfss;
    }
    Fp[0] = L[0] * sound_speed_avg * c_plus;
    Fm[0] = R[0] * sound_speed_avg * c_minus;
    Fp[0] = Fp[0] * mask;
    Fm[0] = Fm[0] * mask;
    Fp[1] = (Fp[0] * L[1]) + (scrD_plus * L[4] * nx);
    Fp[2] = (Fp[0] * L[2]) + (scrD_plus * L[4] * ny);
    Fp[3] = (Fp[0] * L[3]) + (scrD_plus * L[4] * nz);
    Fp[4] = Fp[0] * ((0.5 * (L[1] * L[1] + L[2] * L[2] + L[3] * L[3])) + ((gm / (gm - 1.)) * L[4] / L[0]));
    Fm[1] = (Fm[0] * R[1]) + (scrD_minus * R[4] * nx);
    Fm[2] = (Fm[0] * R[2]) + (scrD_minus * R[4] * ny);
    Fm[3] = (Fm[0] * R[3]) + (scrD_minus * R[4] * nz);
    Fm[4] = Fm[0] * ((0.5 * (R[1] * R[1] + R[2] * R[2] + R[3] * R[3])) + ((gm / (gm - 1.)) * R[4] / R[0]));
    for (int l = 5; l < nv; ++l) { Fp[l] = Fp[0] * L[l]; Fm[l] = Fm[0] * R[l]; }
    for (int l = 0; l < nv; ++l) { Fp[l] = Fp[l] * A; Fm[l] = Fm[l] * A; out[l] = Fp[l] + Fm[l]; }
    return;
  }
  // ausmP / ausmUP / slau share the mass-flux splitting epilogue
  const double rL = L[0], uL = L[1], vL = L[2], wL = L[3], pL = L[4];
  const double rR = R[0], uR = R[1], vR = R[2], wR = R[3], pR = R[4];
  double HL = (0.5 * (uL * uL + vL * vL + wL * wL)) + ((gm / (gm - 1.)) * pL / rL);
  double HR = (0.5 * (uR * uR + vR * vR + wR * wR)) + ((gm / (gm - 1.)) * pR / rR);
  double mass, pbar;
  if (scheme == ORC_AUSMP || scheme == ORC_AUSMUP) {
    double VnL = uL * nx + vL * ny + wL * nz;
    double VnR = uR * nx + vR * ny + wR * nz;
    double cs = std::sqrt(2.0 * (gm - 1.0) * (0.5 * (HL + HR)) / (gm + 1.0));
    double cL, cR;
    if (scheme == ORC_AUSMP) { cL = cs * cs / (std::fmax(cs, std::fabs(VnL))); cR = cs * cs / (std::fmax(cs, std::fabs(VnR))); }
    else { cL = cs * cs / (std::fmax(cs, VnL)); cR = cs * cs / (std::fmax(cs, -VnR)); }
    double C = std::fmin(cL, cR);
    double ML = VnL / C, MR = VnR / C;
    double alfa = 0.1875, fna = 0., Mb = 0.;
    if (scheme == ORC_AUSMUP) {
      Mb = std::sqrt(0.5 * ((VnL * VnL) + (VnR * VnR)) / (C * C));
      double Mo = std::sqrt(std::fmin(1.0, std::fmax(Mb * Mb, MInf * MInf)));
      fna = Mo * (2.0 - Mo);
      alfa = 3.0 * (-4.0 + (5.0 * fna * fna)) / 16.0;
    }
    double alphaL = supersonic_switch(ML), alphaR = supersonic_switch(MR);
    double FmL = (0.5 * (1.0 + fsign(1.0, ML)) * (1.0 - alphaL) * ML) + alphaL * 0.25 * ((1.0 + ML) * (1.0 + ML));
    double betaL = (0.5 * (1.0 + fsign(1.0, ML)) * (1.0 - alphaL)) + alphaL * 0.25 * ((1.0 + ML) * (1.0 + ML)) * (2.0 - ML);
    double FmR = (0.5 * (1.0 - fsign(1.0, MR)) * (1.0 - alphaR) * MR) - alphaR * 0.25 * ((1.0 - MR) * (1.0 - MR));
    double betaR = (0.5 * (1.0 - fsign(1.0, MR)) * (1.0 - alphaR)) + alphaR * 0.25 * ((1.0 - MR) * (1.0 - MR)) * (2.0 + MR);
    double tL = (ML * ML - 1.0), tR = (MR * MR - 1.0);
    FmL = FmL + alphaL * 0.1250 * (tL * tL);
    betaL = betaL + alphaL * alfa * (tL * tL) * ML;
    FmR = FmR - alphaR * 0.1250 * (tR * tR);
    betaR = betaR - alphaR * alfa * (tR * tR) * MR;
    double Mface = FmL + FmR;
    pbar = betaL * pL + betaR * pR;
    if (scheme == ORC_AUSMUP) {
      const double Kp = 0.25, Ku = 0.75, sigma = 1.0;
      double Pu = -Ku * betaL * betaR * (rL + rR) * fna * C * (VnR - VnL);
      double Mp = -2.0 * Kp * std::fmax(1.0 - (sigma * Mb * Mb), 0.0) * (pR - pL) / (fna * (rL + rR) * C * C);
      Mface = FmL + FmR + Mp;
      pbar = betaL * pL + betaR * pR + Pu;
    }
    if (Mface > 0.0) mass = Mface * C * rL; else mass = Mface * C * rR;
  } else {  // slau
    double cL = std::sqrt(gm * pL / rL), cR = std::sqrt(gm * pR / rR);
    double C = 0.5 * (cL + cR);
    double delp = pR - pL;
    double VnL = uL * nx + vL * ny + wL * nz;
    double VnR = uR * nx + vR * ny + wR * nz;
    double ML = VnL / C, MR = VnR / C;
    double alphaL = std::fmax(0.0, 1.0 - std::floor(std::fabs(ML)));
    double alphaR = std::fmax(0.0, 1.0 - std::floor(std::fabs(MR)));
    double betaL = (1.0 - alphaL) * 0.5 * (1.0 + fsign(1.0, ML)) + (alphaL) * 0.25 * (2.0 - ML) * ((ML + 1.0) * (ML + 1.0));
    double betaR = (1.0 - alphaR) * 0.5 * (1.0 - fsign(1.0, MR)) + (alphaR) * 0.25 * (2.0 + MR) * ((MR - 1.0) * (MR - 1.0));
    double vtface = std::sqrt(0.5 * ((uL * uL) + (vL * vL) + (wL * wL) + (uR * uR) + (vR * vR) + (wR * wR)));
    double Mcap = std::fmin(1.0, vtface / C);
    double Xi = (1.0 - Mcap) * (1.0 - Mcap);
    double Vnabs = (rL * std::fabs(VnL) + rR * std::fabs(VnR)) / (rL + rR);
    double fnG = -1.0 * std::fmax(std::fmin(ML, 0.0), -1.0) * std::fmin(std::fmax(MR, 0.0), 1.0);
    pbar = 0.5 * ((pL + pR) + (betaL - betaR) * (pL - pR) + (1.0 - Xi) * (betaL + betaR - 1.0) * (pL + pR));
    double VnabsL = (1.0 - fnG) * Vnabs + fnG * std::fabs(VnL);
    double VnabsR = (1.0 - fnG) * Vnabs + fnG * std::fabs(VnR);
    mass = 0.5 * ((rL * (VnL + VnabsL) + rR * (VnR - VnabsR)) - (Xi * delp / C));
  }
  mass = mass * mask;
  Fp[0] = 0.5 * (mass + std::fabs(mass));
  Fp[1] = Fp[0] * uL; Fp[2] = Fp[0] * vL; Fp[3] = Fp[0] * wL; Fp[4] = Fp[0] * HL;
  Fm[0] = 0.5 * (mass - std::fabs(mass));
  Fm[1] = Fm[0] * uR; Fm[2] = Fm[0] * vR; Fm[3] = Fm[0] * wR; Fm[4] = Fm[0] * HR;
  for (int l = 5; l < nv; ++l) { Fp[l] = Fp[0] * L[l]; Fm[l] = Fm[0] * R[l]; }
  for (int l = 0; l < nv; ++l) out[l] = Fp[l] + Fm[l];
  out[1] = out[1] + (pbar * nx);
  out[2] = out[2] + (pbar * ny);
  out[3] = out[3] + (pbar * nz);
  for (int l = 0; l < nv; ++l) out[l] = out[l] * A;
}

static int flux_dir(Block& B, Arr4& Flux, const Arr4& fl, const Arr4& fr, const Rec4& faces, int ii, int jj, int kk) {
  double L[8], R[8], out[8];
  int nan = 0;
  for (int k = 1; k <= B.kmx - 1 + kk; ++k)
    for (int j = 1; j <= B.jmx - 1 + jj; ++j)
      for (int i = 1; i <= B.imx - 1 + ii; ++i) {
        for (int l = 1; l <= B.nv; ++l) { L[l - 1] = fl(i, j, k, l); R[l - 1] = fr(i, j, k, l); }
        int mask = ii * B.zF[i] + jj * B.zG[j] + kk * B.zH[k];
        flux_kernel(B.c.scheme, B.nv, B.c.gm, B.c.MInf, L, R, faces.A(i, j, k), faces.nx(i, j, k), faces.ny(i, j, k), faces.nz(i, j, k), mask, out);
        for (int l = 1; l <= B.nv; ++l) { Flux(i, j, k, l) = out[l - 1]; if (std::isnan(out[l - 1])) nan = 1; }
      }
  return nan;
}

// e.g. ausm.f90:160-216 compute_fluxes: F, G, then H = 0 when kmx == 2; NaN -> Fatal_error
void Block::compute_fluxes() {
  if (flux_dir(*this, F, xl, xr, If, 1, 0, 0)) error |= 1;
  if (flux_dir(*this, G, yl, yr, Jf, 0, 1, 0)) error |= 1;
  if (kmx == 2) std::fill(H.d.begin(), H.d.end(), 0.0);
  else if (flux_dir(*this, H, zl, zr, Kf, 0, 0, 1)) error |= 1;
}

// scheme.f90:111-141
void Block::compute_residue() {
  for (int l = 1; l <= nv; ++l)
    for (int k = 1; k <= kmx - 1; ++k)
      for (int j = 1; j <= jmx - 1; ++j)
        for (int i = 1; i <= imx - 1; ++i)
          residue(i, j, k, l) = (F(i + 1, j, k, l) - F(i, j, k, l)) + (G(i, j + 1, k, l) - G(i, j, k, l)) + (H(i, j, k + 1, l) - H(i, j, k, l));
}

// interface1.f90:96-493: pack 3 interior layers (order n -> l -> outer transverse -> inner transverse)
void Block::pack(int face) {
  Frame f = make_frame(*this, face);
  std::vector<double>& buf = sendbuf[face - 1];
  buf.resize((size_t)f.na * f.nb * nv * 3);
  size_t count = 0;
  for (int n = 1; n <= nv; ++n)
    for (int l = 1; l <= 3; ++l)
      for (int b = 1; b <= f.nb; ++b)
        for (int a = 1; a <= f.na; ++a) {
          int i, j, k; f.interior(l, a, b, i, j, k);
          buf[count++] = qp(i, j, k, n);
        }
}

// unpack honouring Pxlo..Pxhi step PxDir and dir_switch (interface1.f90:144-168)
void Block::unpack(int face) {
  Frame f = make_frame(*this, face);
  const std::vector<double>& buf = recvbuf[face - 1];
  const int fi = face - 1;
  const int alo = c.plo[fi][0], ahi = c.phi[fi][0], ad = c.pdir[fi][0];
  const int blo = c.plo[fi][1], bhi = c.phi[fi][1], bd = c.pdir[fi][1];
  size_t count = 0;
  for (int n = 1; n <= nv; ++n)
    for (int l = 1; l <= 3; ++l) {
      if (c.dir_switch[fi] == 0) {
        for (int b = blo; (bd > 0 ? b <= bhi : b >= bhi); b += bd)
          for (int a = alo; (ad > 0 ? a <= ahi : a >= ahi); a += ad) {
            int i, j, k; f.ghost(l, a, b, i, j, k);
            qp(i, j, k, n) = buf[count++];
          }
      } else {
        for (int a = alo; (ad > 0 ? a <= ahi : a >= ahi); a += ad)
          for (int b = blo; (bd > 0 ? b <= bhi : b >= bhi); b += bd) {
            int i, j, k; f.ghost(l, a, b, i, j, k);
            qp(i, j, k, n) = buf[count++];
          }
      }
    }
}

}  // namespace orc
