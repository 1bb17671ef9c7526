// TEST INFRASTRUCTURE ONLY -- CPU oracle for the FEST-3D explicit residual + update path.
// Literal re-statement of the reference Fortran (file:line cited per function).  PARITY UNPINNED
// against a runnable reference (no Fortran compiler here); pinned on the reference's unit-test KATs.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <vector>
#include "oracle_abi.h"

namespace orc {

// Fortran-ordered arrays with explicit lower bounds (first index fastest).
struct Arr3 {
  int l0 = 0, l1 = 0, l2 = 0, n0 = 0, n1 = 0, n2 = 0;
  std::vector<double> d;
  void alloc(int a0, int b0, int a1, int b1, int a2, int b2, double v = 0.0) {
    l0 = a0; l1 = a1; l2 = a2; n0 = b0 - a0 + 1; n1 = b1 - a1 + 1; n2 = b2 - a2 + 1;
    d.assign((size_t)n0 * n1 * n2, v);
  }
  inline size_t idx(int i, int j, int k) const {
    return (size_t)(i - l0) + (size_t)n0 * ((size_t)(j - l1) + (size_t)n1 * (size_t)(k - l2));
  }
  inline double& operator()(int i, int j, int k) { return d[idx(i, j, k)]; }
  inline double operator()(int i, int j, int k) const { return d[idx(i, j, k)]; }
  size_t size() const { return d.size(); }
};

struct Arr4 {
  int l0 = 0, l1 = 0, l2 = 0, n0 = 0, n1 = 0, n2 = 0, nv = 0;  // 4th index runs 1..nv
  std::vector<double> d;
  void alloc(int a0, int b0, int a1, int b1, int a2, int b2, int nvar, double v = 0.0) {
    l0 = a0; l1 = a1; l2 = a2; n0 = b0 - a0 + 1; n1 = b1 - a1 + 1; n2 = b2 - a2 + 1; nv = nvar;
    d.assign((size_t)n0 * n1 * n2 * nv, v);
  }
  inline size_t idx(int i, int j, int k, int l) const {
    return (size_t)(i - l0) + (size_t)n0 * ((size_t)(j - l1) + (size_t)n1 * ((size_t)(k - l2) + (size_t)n2 * (size_t)(l - 1)));
  }
  inline double& operator()(int i, int j, int k, int l) { return d[idx(i, j, k, l)]; }
  inline double operator()(int i, int j, int k, int l) const { return d[idx(i, j, k, l)]; }
  size_t size() const { return d.size(); }
};

// AoS array of 4-double records, lower bound -2 on all axes (vartypes.f90:29-46).
struct Rec4 {
  int n0 = 0, n1 = 0, n2 = 0;
  std::vector<double> d;
  void alloc(int hi0, int hi1, int hi2) {
    n0 = hi0 + 3; n1 = hi1 + 3; n2 = hi2 + 3;
    d.assign((size_t)4 * n0 * n1 * n2, 0.0);
  }
  inline const double* at(int i, int j, int k) const {
    return &d[4 * ((size_t)(i + 2) + (size_t)n0 * ((size_t)(j + 2) + (size_t)n1 * (size_t)(k + 2)))];
  }
  // facetype {A,nx,ny,nz}
  inline double A(int i, int j, int k) const { return at(i, j, k)[0]; }
  inline double nx(int i, int j, int k) const { return at(i, j, k)[1]; }
  inline double ny(int i, int j, int k) const { return at(i, j, k)[2]; }
  inline double nz(int i, int j, int k) const { return at(i, j, k)[3]; }
  // celltype {volume,centerx,centery,centerz}
  inline double vol(int i, int j, int k) const { return at(i, j, k)[0]; }
  inline double cx(int i, int j, int k) const { return at(i, j, k)[1]; }
  inline double cy(int i, int j, int k) const { return at(i, j, k)[2]; }
  inline double cz(int i, int j, int k) const { return at(i, j, k)[3]; }
};

// SST constants (global_sst.f90:6-16)
constexpr double sigma_k1 = 0.85, sigma_k2 = 1.0, sigma_w1 = 0.5, sigma_w2 = 0.856;
constexpr double beta1 = 0.075, beta2 = 0.0828, bstar = 0.09, kappa_sst = 0.41, a1_sst = 0.31;

struct Block {
  OracleConfig c;
  int imx, jmx, kmx, nv, n_grad;
  Arr4 qp, U_store, R_store, residue, F, G, H;
  Arr4 xl, xr, yl, yr, zl, zr;        // x_qp_left ... (face_interpolant.f90:43-55)
  Arr4 gx, gy, gz;                    // gradqp_x/y/z (0:imx,0:jmx,0:kmx,n_grad)
  Arr3 Temp, delta_t, mu, mu_t, F1, dist, pdif;
  Arr3 dvdy;                          // lctm2015: DCCVn . CCnormal of CC.f90 (a function of the wall distance only, see add_sst_source_lctm2015)
  bool dvdy_ready = false;
  Rec4 cells, If, Jf, Kf;
  std::vector<int> zF, zG, zH;        // make_{F,G,H}_flux_zero, 1-based (bc.f90:53-66)
  double c1, c2, c3;                  // bc.f90:48-50
  double gama1, gama2;                // global_sst.f90:15-16 (module variables, rewritten by sst2003)
  int ppm_flag = 0;                   // sticky (boundary_state_reconstruction.f90:19)
  int current_iter = 1;
  int error = 0;
  std::vector<double> sendbuf[6], recvbuf[6];
  double res_abs_local[16];

  void setup(const OracleConfig& cfg);
  // update.f90
  void refresh_temp();
  void total_residue();               // everything after apply_interface
  void compute_time_step();
  void update_with(double TF, double SF, bool TU, bool have_store);
  int update_with_lusgs();            // lusgs.f90:134-183 (oracle_lusgs.cpp); 64 = a model whose LU-SGS routine is not restated
  void absolute_resnorm();
  // pieces
  void populate_ghost_primitive();
  void compute_face_interpolant();
  void reconstruct_boundary_state();
  void compute_fluxes();
  void evaluate_all_gradients();
  void calculate_viscosity();
  void compute_viscous_fluxes();
  void compute_residue();
  void add_source_term_residue();
  void pack(int face);
  void unpack(int face);
};

// stand-alone kernels
void flux_kernel(int scheme, int nv, double gm, double MInf, const double* L, const double* R,
                 double A, double nx, double ny, double nz, int mask, double* out);

}  // namespace orc

struct OracleWorld {
  std::vector<orc::Block> blocks;
};
