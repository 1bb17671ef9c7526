// TEST INFRASTRUCTURE ONLY -- CPU oracle, part 3: the iteration driver (update.f90:129-226 get_next_solution,
// interface1.f90 halo swap between the blocks of one world, resnorm.f90:201-225) and the C ABI.
// PARITY UNPINNED (see oracle_abi.h).
#include <thread>
#include "oracle_core.hpp"

using orc::Block;

namespace {

// interface1.f90:96-493 apply_interface for all blocks at once.  Every exchange reads interior cells
// only, so packing everything first and unpacking afterwards equals the pairwise MPI_SENDRECVs.
void apply_interface(OracleWorld* w) {
  for (Block& B : w->blocks)
    for (int face = 1; face <= 6; ++face)
      if (B.c.bc_id[face - 1] >= 0 || B.c.pbc_id[face - 1] >= 0) B.pack(face);
  for (Block& B : w->blocks)
    for (int face = 1; face <= 6; ++face) {
      int nb = B.c.bc_id[face - 1] >= 0 ? B.c.bc_id[face - 1] : B.c.pbc_id[face - 1];
      if (nb < 0 || nb >= (int)w->blocks.size()) continue;   // neighbour not in this world: ghosts keep their values
      // the neighbour's send buffer for the face that is attached to this one
      int of = B.c.otherface[face - 1];
      B.recvbuf[face - 1] = w->blocks[nb].sendbuf[of - 1];
      B.unpack(face);
    }
}

template <class Fn>
void for_blocks(OracleWorld* w, Fn fn) {
  if (w->blocks.size() == 1) { fn(w->blocks[0]); return; }
  std::vector<std::thread> th;
  for (Block& B : w->blocks) th.emplace_back([&B, &fn]() { fn(B); });   // one "rank" per block
  for (auto& t : th) t.join();
}

void residual_all(OracleWorld* w) {
  apply_interface(w);
  for_blocks(w, [](Block& B) { B.total_residue(); });
}

void blend(Block& B, double a, double b) {   // qp = a*U_store + b*qp over the whole array (update.f90:203,206,214)
  for (size_t n = 0; n < B.qp.d.size(); ++n) B.qp.d[n] = a * B.U_store.d[n] + b * B.qp.d[n];
}

}  // namespace

extern "C" {

OracleWorld* oracle_create(int n_blocks, const OracleConfig* cfgs) {
  OracleWorld* w = new OracleWorld();
  w->blocks.resize(n_blocks);
  for (int b = 0; b < n_blocks; ++b) w->blocks[b].setup(cfgs[b]);
  return w;
}

void oracle_destroy(OracleWorld* w) { delete w; }

int oracle_set_geometry(OracleWorld* w, int b, const double* cells, const double* Ifaces, const double* Jfaces,
                        const double* Kfaces, const double* dist) {
  Block& B = w->blocks[b];
  std::memcpy(B.cells.d.data(), cells, B.cells.d.size() * sizeof(double));
  std::memcpy(B.If.d.data(), Ifaces, B.If.d.size() * sizeof(double));
  std::memcpy(B.Jf.d.data(), Jfaces, B.Jf.d.size() * sizeof(double));
  std::memcpy(B.Kf.d.data(), Kfaces, B.Kf.d.size() * sizeof(double));
  if (dist && B.dist.size()) std::memcpy(B.dist.d.data(), dist, B.dist.d.size() * sizeof(double));
  return 0;
}

int oracle_set_state(OracleWorld* w, int b, const double* qp) {
  Block& B = w->blocks[b];
  std::memcpy(B.qp.d.data(), qp, B.qp.d.size() * sizeof(double));
  return 0;
}

int oracle_get_state(OracleWorld* w, int b, double* qp) {
  Block& B = w->blocks[b];
  std::memcpy(qp, B.qp.d.data(), B.qp.d.size() * sizeof(double));
  return 0;
}

int oracle_residual(OracleWorld* w, int current_iter) {
  for (Block& B : w->blocks) { B.current_iter = current_iter; B.refresh_temp(); }
  residual_all(w);
  int e = 0;
  for (Block& B : w->blocks) e |= B.error;
  return e;
}

int oracle_get_residue(OracleWorld* w, int b, double* residue) {
  Block& B = w->blocks[b];
  std::memcpy(residue, B.residue.d.data(), B.residue.d.size() * sizeof(double));
  return 0;
}

// update.f90:129-226 + resnorm.f90:62-91
int oracle_step(OracleWorld* w, int current_iter, double* res_abs) {
  for (Block& B : w->blocks) { B.current_iter = current_iter; B.refresh_temp(); }
  const int ta = w->blocks[0].c.time_accuracy;
  auto upd = [w](double TF, double SF, bool TU, bool store) { for_blocks(w, [=](Block& B) { B.update_with(TF, SF, TU, store); }); };
  auto dt = [w]() { for_blocks(w, [](Block& B) { B.compute_time_step(); }); };
  switch (ta) {
    case ORC_T_NONE:
      residual_all(w); dt(); upd(1., 1., false, false);
      break;
    case ORC_T_RK4:
      for (Block& B : w->blocks) { std::fill(B.R_store.d.begin(), B.R_store.d.end(), 0.0); B.U_store.d = B.qp.d; }
      residual_all(w); dt(); upd(0.5, 1., false, true);
      residual_all(w); upd(0.5, 2., false, true);
      residual_all(w); upd(1.0, 2., false, true);
      residual_all(w); upd(1. / 6., 1., true, true);
      break;
    case ORC_T_RK2:
      for (Block& B : w->blocks) { std::fill(B.R_store.d.begin(), B.R_store.d.end(), 0.0); B.U_store.d = B.qp.d; }
      residual_all(w); dt(); upd(0.5, 1., false, true);
      residual_all(w); upd(0.5, 1., true, true);
      break;
    case ORC_T_TVDRK3:
      for (Block& B : w->blocks) B.U_store.d = B.qp.d;
      residual_all(w); dt(); upd(1.0, 1., false, false);
      residual_all(w); upd(1.0, 1., false, false);
      for (Block& B : w->blocks) blend(B, 0.75, 0.25);
      residual_all(w); upd(1.0, 1., false, false);
      for (Block& B : w->blocks) blend(B, (1. / 3.), (2. / 3.));
      break;
    case ORC_T_TVDRK2:
      for (Block& B : w->blocks) B.U_store.d = B.qp.d;
      residual_all(w); dt(); upd(1.0, 1., false, false);
      residual_all(w); upd(1.0, 1., false, false);
      for (Block& B : w->blocks) blend(B, 0.5, 0.5);
      break;
    case ORC_T_IMPLICIT: {   // update.f90:216-219
      residual_all(w); dt();
      int rc = 0;
      for (Block& B : w->blocks) rc |= B.update_with_lusgs();
      if (rc) return rc;
      break;
    }
    default:
      return 64;  // plusgs: not on this path
  }
  // find_resnorm: per-block sums, "allgather", sum over blocks, sqrt / abs (resnorm.f90:171-225)
  const int nv = w->blocks[0].nv;
  for_blocks(w, [](Block& B) { B.absolute_resnorm(); });
  for (int l = 0; l <= nv; ++l) {
    double s = 0.;
    for (Block& B : w->blocks) s = s + B.res_abs_local[l];
    res_abs[l] = (l == 0) ? std::fabs(s) : std::sqrt(s);
  }
  int e = 0;
  for (Block& B : w->blocks) e |= B.error;
  return e;
}

int oracle_get_aux(OracleWorld* w, int b, int which, double* out) {
  Block& B = w->blocks[b];
  const std::vector<double>* src = nullptr;
  switch (which) {
    case 0: src = &B.delta_t.d; break;
    case 1: src = &B.mu.d; break;
    case 2: src = &B.mu_t.d; break;
    case 3: src = &B.F1.d; break;
    case 4: src = &B.Temp.d; break;
    case 5: src = &B.dvdy.d; break;   // lctm2015 only
    case 10: src = &B.xl.d; break; case 11: src = &B.xr.d; break;
    case 12: src = &B.yl.d; break; case 13: src = &B.yr.d; break;
    case 14: src = &B.zl.d; break; case 15: src = &B.zr.d; break;
    case 20: src = &B.F.d; break; case 21: src = &B.G.d; break; case 22: src = &B.H.d; break;
    case 30: src = &B.gx.d; break; case 31: src = &B.gy.d; break; case 32: src = &B.gz.d; break;
    default: return 1;
  }
  std::memcpy(out, src->data(), src->size() * sizeof(double));
  return 0;
}

void oracle_kat_flux(int scheme, int n_var, double gm, double MInf, const double* left, const double* right,
                     const double* face, int mask, double* flux) {
  orc::flux_kernel(scheme, n_var, gm, MInf, left, right, face[0], face[1], face[2], face[3], mask, flux);
}

// 1-D line of n cells (Fortran cells -2..n-3 i.e. imx = n-5 ... the reference tests use cells -2..7, imx=5):
// runs the block machinery on a (imx,2,2) block with the line along i.
void oracle_kat_states(int interpolant, int n, const double* q, const double* vol, int limiter, double* left, double* right) {
  OracleConfig c;
  std::memset(&c, 0, sizeof(c));
  c.imx = n - 5; c.jmx = 2; c.kmx = 2; c.n_var = 5;
  c.interpolant = interpolant; c.scheme = ORC_AUSM;
  for (int d = 0; d < 3; ++d) { c.limiter[d] = limiter; c.tlimiter[d] = limiter; }
  for (int f = 0; f < 6; ++f) { c.bc_id[f] = 0; c.pbc_id[f] = -1; }
  Block B; B.setup(c);
  for (int l = 1; l <= 5; ++l)
    for (int k = -2; k <= 4; ++k) for (int j = -2; j <= 4; ++j) for (int i = -2; i <= c.imx + 2; ++i) B.qp(i, j, k, l) = q[i + 2];
  for (int k = -2; k <= 4; ++k) for (int j = -2; j <= 4; ++j) for (int i = -2; i <= c.imx + 2; ++i)
    B.cells.d[4 * ((size_t)(i + 2) + (size_t)B.cells.n0 * ((size_t)(j + 2) + (size_t)B.cells.n1 * (size_t)(k + 2)))] = vol ? vol[i + 2] : 1.0;
  B.compute_face_interpolant();
  for (int i = 0; i <= c.imx + 1; ++i) { left[i] = B.xl(i, 1, 1, 1); right[i] = B.xr(i, 1, 1, 1); }
}

// test_residue.f90:19-30: compute_residue (scheme.f90:111-141) on hand-set flux arrays.  F is (1:imx,1:jmx-1,1:kmx-1,nv), G and H
// alike (Fortran order, i fastest); residue is (1:imx-1,1:jmx-1,1:kmx-1,nv).  Also returns the boundary mass-flux imbalance of
// get_absolute_resnorm (resnorm.f90:190-198) in *merror.
void oracle_kat_residue(int imx, int jmx, int kmx, int n_var, const double* F, const double* G, const double* H, double* residue, double* merror) {
  OracleConfig c;
  std::memset(&c, 0, sizeof(c));
  c.imx = imx; c.jmx = jmx; c.kmx = kmx; c.n_var = n_var; c.scheme = ORC_AUSM; c.interpolant = ORC_NONE;
  c.density_inf = 1.0; c.vel_mag = 1.0; c.pressure_inf = 1.0; c.gm = 1.4;
  for (int f = 0; f < 6; ++f) { c.bc_id[f] = 0; c.pbc_id[f] = -1; }
  Block B; B.setup(c);
  std::memcpy(B.F.d.data(), F, B.F.d.size() * sizeof(double));
  std::memcpy(B.G.d.data(), G, B.G.d.size() * sizeof(double));
  std::memcpy(B.H.d.data(), H, B.H.d.size() * sizeof(double));
  B.compute_residue();
  std::memcpy(residue, B.residue.d.data(), B.residue.d.size() * sizeof(double));
  B.absolute_resnorm();
  if (merror) *merror = B.res_abs_local[0];
}

// wall_dist.f90:84-131 find_wall_dist: brute-force minimum distance of every node (-2:imx+3, ...) to the wall surface nodes, then the
// cell value as 0.125 x the sum of its eight nodes in the reference's order.  nodes: (x,y,z) records, i fastest; wall: n x 3.
void oracle_find_wall_dist(int imx, int jmx, int kmx, const double* nodes, const double* wall, long long n_wall, double* dist_out) {
  const long long ni = imx + 6, nj = jmx + 6, nk = kmx + 6;
  std::vector<double> nd((size_t)(ni * nj * nk));
  for (long long n = 0; n < ni * nj * nk; ++n) {
    double best = 1.e+20;
    const double x = nodes[3 * n], y = nodes[3 * n + 1], z = nodes[3 * n + 2];
    for (long long w = 0; w < n_wall; ++w) {
      const double a = wall[3 * w] - x, b = wall[3 * w + 1] - y, c = wall[3 * w + 2] - z;
      const double current_dist = std::sqrt((a * a) + (b * b) + (c * c));
      best = std::fmin(best, current_dist);
    }
    nd[(size_t)n] = best;
  }
  auto at = [&](int i, int j, int k) { return nd[(size_t)((i + 2) + ni * ((j + 2) + nj * (long long)(k + 2)))]; };
  for (int k = -2; k <= kmx + 2; ++k)
    for (int j = -2; j <= jmx + 2; ++j)
      for (int i = -2; i <= imx + 2; ++i)
        dist_out[(size_t)((i + 2) + (long long)(imx + 5) * ((j + 2) + (long long)(jmx + 5) * (k + 2)))] =
            0.125 * (at(i, j, k) + at(i, j + 1, k) + at(i, j + 1, k + 1) + at(i, j, k + 1) + at(i + 1, j, k + 1) + at(i + 1, j, k) + at(i + 1, j + 1, k) +
                     at(i + 1, j + 1, k + 1));
}

}  // extern "C"
