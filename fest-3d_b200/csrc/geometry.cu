// Grid and metric set-up on the device -- SURVEY 8(f) rank 2.  Replaces, for a block whose grid file the host has read,
//   ghost_grid                      src/grid.f90:137-236      ghost nodes by linear extrapolation, i then j then k
//   compute_face_area_vectors       src/geometry.f90:236-314  half cross product of the face diagonals
//   compute_face_areas              src/geometry.f90:216-233  A = |vector|, A = 0 on the four face layers of a pole (-7)
//   normalize_face_normals          src/geometry.f90:43-213   n = vector / A where A /= 0; pole layers copy the normal of face 2 / mx-1
//   compute_volumes                 src/geometry.f90:316-496  two 5-tetrahedra splits of the hexahedron, the larger one; cells 0..imx
//   compute_cell_centre             src/geometry.f90:500-545  0.125 x the sum of the eight nodes in the reference's order
// and the 16-array AoS host->device upload of fest3d_gpu_set_geometry: the metrics are written straight into the SoA fields the
// kernels read (ctx.hpp: G_VOL .. G_KNZ), the face records of the ghost-gradient rule (gradients.f90:549-612, read by the reference
// through an Ifaces-shaped dummy) are gathered on the device with the same linear offset.
//
// This file is compiled with -fmad=false: every operation is then the IEEE operation the reference's statement performs
// (subtract, multiply, divide and sqrt are correctly rounded on the device), so the result equals the host computation bit for
// bit and the parity test can ask for equality.
#include "ctx.hpp"

namespace f3d {

struct P3 { double x, y, z; };
__device__ __forceinline__ P3 operator-(const P3& a, const P3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }

struct NodeView {   // nodes(-2:imx+3,-2:jmx+3,-2:kmx+3) of {x,y,z}
  double* p; int ni, nj, nk;
  __device__ __forceinline__ long long at(int i, int j, int k) const { return 3 * ((i + 2) + (long long)ni * ((j + 2) + (long long)nj * (k + 2))); }
  __device__ __forceinline__ P3 get(int i, int j, int k) const { const double* q = p + at(i, j, k); return {q[0], q[1], q[2]}; }
};

// interior nodes (1:imx,1:jmx,1:kmx), the body of the grid file (grid.f90:78-133), into the ghosted array
__global__ void k_place_nodes(NodeView nv, const double* __restrict__ grid, int imx, int jmx, int kmx) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x, tot = (long long)imx * jmx * kmx;
  if (n >= tot) return;
  const int i = (int)(n % imx), j = (int)((n / imx) % jmx), k = (int)(n / ((long long)imx * jmx));
  double* q = nv.p + nv.at(i + 1, j + 1, k + 1);
  q[0] = grid[3 * n]; q[1] = grid[3 * n + 1]; q[2] = grid[3 * n + 2];
}

// one pass of ghost_grid along `axis` (node count mx): every line of the WHOLE array (ghost lines of the other axes included,
// like the reference's (:,:) sections) is extended by three nodes at either end, each from the two before it
__global__ void k_extend_nodes(NodeView nv, int axis, int mx) {
  const int na = (axis == 0) ? nv.nj : nv.ni, nb = (axis == 2) ? nv.nj : nv.nk;
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (a >= na || b >= nb) return;
  auto ptr = [&](int t) {
    int i, j, k;
    if (axis == 0) { i = t; j = a - 2; k = b - 2; } else if (axis == 1) { i = a - 2; j = t; k = b - 2; } else { i = a - 2; j = b - 2; k = t; }
    return nv.p + nv.at(i, j, k);
  };
  for (int c = 0; c < 3; ++c) {
    double x2 = ptr(2)[c], x1 = ptr(1)[c];
    for (int t = 0; t >= -2; --t) { const double x0 = 2 * x1 - x2; ptr(t)[c] = x0; x2 = x1; x1 = x0; }
    x2 = ptr(mx - 1)[c]; x1 = ptr(mx)[c];
    for (int t = mx + 1; t <= mx + 3; ++t) { const double x0 = 2 * x1 - x2; ptr(t)[c] = x0; x2 = x1; x1 = x0; }
  }
}

__device__ __forceinline__ P3 half_cross(const P3& d1, const P3& d2) {
  return {0.5 * (d1.y * d2.z - d1.z * d2.y), 0.5 * (d1.z * d2.x - d1.x * d2.z), 0.5 * (d1.x * d2.y - d1.y * d2.x)};
}

// area vector of face (i,j,k) of direction d (geometry.f90:261-303)
__device__ __forceinline__ P3 face_vector(const NodeView& nv, int d, int i, int j, int k) {
  if (d == 0) return half_cross(nv.get(i, j + 1, k + 1) - nv.get(i, j, k), nv.get(i, j, k + 1) - nv.get(i, j + 1, k));
  if (d == 1) return half_cross(nv.get(i + 1, j, k + 1) - nv.get(i, j, k), nv.get(i + 1, j, k) - nv.get(i, j, k + 1));
  return half_cross(nv.get(i + 1, j + 1, k) - nv.get(i, j, k), nv.get(i, j + 1, k) - nv.get(i + 1, j, k));
}

__device__ __forceinline__ void unit_normal(P3 v, double& A, P3& n) {
  A = sqrt(((v.x * v.x) + (v.y * v.y)) + (v.z * v.z));
  n = v;
  if (A != 0.) { n.x = v.x / A; n.y = v.y / A; n.z = v.z / A; }
}

__device__ __forceinline__ double tet(const P3& p1, const P3& p2, const P3& p3, const P3& p4) {   // geometry.f90:316-346
  const P3 a = p2 - p1, b = p3 - p1, c = p4 - p1;
  const double v = ((c.x * ((a.y * b.z) - (a.z * b.y))) + (c.y * ((a.z * b.x) - (a.x * b.z)))) + (c.z * ((a.x * b.y) - (a.y * b.x)));
  return (-v) / 6.0;
}

// one thread per index triple of -2..imx+3 x -2..jmx+3 x -2..kmx+3: the three faces and the cell that carry this index
__global__ void k_metrics(const Layout L, NodeView nv, double* __restrict__ geom, int b0, int b1, int b2, int b3, int b4, int b5, int* __restrict__ err) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) - 2, j = (int)(blockIdx.y * blockDim.y + threadIdx.y) - 2, k = (int)blockIdx.z - 2;
  if (i > L.imx + 3 || j > L.jmx + 3) return;
  const long long fs = L.fs, c = L.idx(i, j, k);
  const int bc[6] = {b0, b1, b2, b3, b4, b5};
  const int mx[3] = {L.imx, L.jmx, L.kmx};
  const int ix[3] = {i, j, k};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    bool in = true;   // the face array of direction d extends one further along d only
#pragma unroll
    for (int e = 0; e < 3; ++e) in = in && (ix[e] <= mx[e] + (e == d ? 3 : 2));
    if (!in) continue;
    double A; P3 n;
    unit_normal(face_vector(nv, d, i, j, k), A, n);
    const bool pole_lo = bc[2 * d] == -7 && ix[d] <= 1, pole_hi = bc[2 * d + 1] == -7 && ix[d] >= mx[d];
    if (pole_lo || pole_hi) {   // A = 0 and the normal of face 2 / mx-1 of the same line
      int s[3] = {i, j, k};
      s[d] = pole_lo ? 2 : mx[d] - 1;
      double As;
      unit_normal(face_vector(nv, d, s[0], s[1], s[2]), As, n);
      A = 0.;
    }
    double* g = geom + (long long)(G_IA + 4 * d) * fs + c;
    g[0] = A; g[fs] = n.x; g[2 * fs] = n.y; g[3 * fs] = n.z;
  }
  if (i <= L.imx + 2 && j <= L.jmx + 2 && k <= L.kmx + 2) {
    const P3 p000 = nv.get(i, j, k), p100 = nv.get(i + 1, j, k), p110 = nv.get(i + 1, j + 1, k), p010 = nv.get(i, j + 1, k);
    const P3 p001 = nv.get(i, j, k + 1), p101 = nv.get(i + 1, j, k + 1), p111 = nv.get(i + 1, j + 1, k + 1), p011 = nv.get(i, j + 1, k + 1);
    // centre: the reference's summation order (geometry.f90:509-517)
    const double cx = 0.125 * (((((((p000.x + p100.x) + p110.x) + p111.x) + p101.x) + p010.x) + p011.x) + p001.x);
    const double cy = 0.125 * (((((((p000.y + p100.y) + p110.y) + p111.y) + p101.y) + p010.y) + p011.y) + p001.y);
    const double cz = 0.125 * (((((((p000.z + p100.z) + p110.z) + p111.z) + p101.z) + p010.z) + p011.z) + p001.z);
    double vol = 1.0;   // cells outside 0..imx keep 1 (geometry.f90:448)
    if (i >= 0 && i <= L.imx && j >= 0 && j <= L.jmx && k >= 0 && k <= L.kmx) {
      // p1..p8 of the reference: (i,j,k) (i+1,j,k) (i+1,j+1,k) (i,j+1,k) (i,j,k+1) (i+1,j,k+1) (i+1,j+1,k+1) (i,j+1,k+1)
      const P3 &p1 = p000, &p2 = p100, &p3 = p110, &p4 = p010, &p5 = p001, &p6 = p101, &p7 = p111, &p8 = p011;
      double v1 = tet(p1, p5, p8, p6);
      v1 = v1 + tet(p7, p8, p6, p3);
      v1 = v1 + tet(p8, p4, p1, p3);
      v1 = v1 + tet(p6, p1, p3, p8);
      v1 = v1 + tet(p1, p2, p6, p3);
      double v2 = tet(p2, p6, p5, p7);
      v2 = v2 + tet(p8, p5, p7, p4);
      v2 = v2 + tet(p5, p1, p2, p4);
      v2 = v2 + tet(p7, p2, p4, p5);
      v2 = v2 + tet(p2, p3, p7, p4);
      vol = fmax(v2, v1);
      if (!(vol > 0.)) {   // Fatal_error of geometry.f90:476-494
        const int old = atomicOr(&err[0], F3D_ERR_GEOMETRY);
        if ((old & F3D_ERR_GEOMETRY) == 0) { err[1] = i; err[2] = j; err[3] = k; }
      }
    }
    geom[(long long)G_VOL * fs + c] = vol;
    geom[(long long)G_CX * fs + c] = cx; geom[(long long)G_CY * fs + c] = cy; geom[(long long)G_CZ * fs + c] = cz;
  }
}

// Face records of the ghost-gradient rule.  apply_gradient_bc_face declares its face argument with the Ifaces shape
// (-2:imx+3,-2:jmx+2,-2:kmx+2) and is handed Jfaces / Kfaces as well (gradients.f90:549-612): element (i,j,k) is then the record
// at linear offset (i+2) + (imx+6)*((j+2) + (jmx+5)*(k+2)) of the ACTUAL array, whose own shape decodes that offset to a
// different face.  Same gather as the host loop of fest3d_gpu_set_geometry, one thread per face cell.
__global__ void k_gather_gbc(const Layout L, const double* __restrict__ geom, double* __restrict__ gbc, int f, long long off_f) {
  const int ax = f / 2, a_ax = (ax == 0) ? 1 : 0, b_ax = (ax == 2) ? 1 : 2;
  const int mx[3] = {L.imx, L.jmx, L.kmx};
  const int na = mx[a_ax] - 1, nb = mx[b_ax] - 1;
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (a >= na || b >= nb) return;
  int idx[3]; idx[a_ax] = a + 1; idx[b_ax] = b + 1; idx[ax] = (f % 2 == 0) ? 1 : mx[ax];
  const long long off = (idx[0] + 2) + (long long)(L.imx + 6) * ((idx[1] + 2) + (long long)(L.jmx + 5) * (idx[2] + 2));
  const long long n0 = L.imx + (ax == 0 ? 6 : 5), n1 = L.jmx + (ax == 1 ? 6 : 5);   // extents of the actual array
  const int ii = (int)(off % n0) - 2, jj = (int)((off / n0) % n1) - 2, kk = (int)(off / (n0 * n1)) - 2;
  const double* g = geom + (long long)(G_IA + 4 * ax) * L.fs + L.idx(ii, jj, kk);
  double* o = gbc + off_f + 4 * ((long long)b * na + a);
  o[0] = g[0]; o[1] = g[L.fs]; o[2] = g[2 * L.fs]; o[3] = g[3 * L.fs];
}

// AoS records (reference layout) back from the SoA fields: 4 fields -> n0 x n1 x n2 records of 4 doubles
__global__ void k_records_out(const Layout L, const double* __restrict__ field0, double* __restrict__ out, int n0, int n1, int n2) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y, c = blockIdx.z;
  if (a >= n0) return;
  const long long s = L.idx(a - 2, b - 2, c - 2);
  double* o = out + 4 * (a + (long long)n0 * (b + (long long)n1 * c));
  o[0] = field0[s]; o[1] = field0[L.fs + s]; o[2] = field0[2 * L.fs + s]; o[3] = field0[3 * L.fs + s];
}

int launch_setup_geometry(Ctx* ctx, const double* grid_host, double* nodes_out) {
  const Layout& L = ctx->P.L;
  NodeView nv{nullptr, L.imx + 6, L.jmx + 6, L.kmx + 6};
  const long long n_nodes = (long long)nv.ni * nv.nj * nv.nk, n_int = (long long)L.imx * L.jmx * L.kmx;
  double* d_grid = nullptr;
  F3D_CUDA(cudaMalloc((void**)&nv.p, sizeof(double) * 3 * n_nodes));
  F3D_CUDA(cudaMalloc((void**)&d_grid, sizeof(double) * 3 * n_int));
  F3D_CUDA(cudaMemsetAsync(nv.p, 0, sizeof(double) * 3 * n_nodes, ctx->stream));
  F3D_CUDA(cudaMemcpyAsync(d_grid, grid_host, sizeof(double) * 3 * n_int, cudaMemcpyHostToDevice, ctx->stream));
  k_place_nodes<<<(unsigned)((n_int + 255) / 256), 256, 0, ctx->stream>>>(nv, d_grid, L.imx, L.jmx, L.kmx);
  const int mx[3] = {L.imx, L.jmx, L.kmx};
  for (int axis = 0; axis < 3; ++axis) {
    const int na = (axis == 0) ? nv.nj : nv.ni, nb = (axis == 2) ? nv.nj : nv.nk;
    k_extend_nodes<<<dim3((na + 63) / 64, nb), 64, 0, ctx->stream>>>(nv, axis, mx[axis]);
  }
  const int* b = ctx->cfg.bc_id;
  k_metrics<<<dim3((L.imx + 6 + 31) / 32, (L.jmx + 6 + 3) / 4, L.kmx + 6), dim3(32, 4), 0, ctx->stream>>>(L, nv, ctx->geom, b[0], b[1], b[2], b[3], b[4], b[5],
                                                                                                     ctx->err_dev);
  ctx->launches += 5;
  if (ctx->P.viscous) {
    for (int f = 0; f < 6; ++f) {
      const int ax = f / 2, a_ax = (ax == 0) ? 1 : 0, b_ax = (ax == 2) ? 1 : 2;
      k_gather_gbc<<<dim3((mx[a_ax] - 1 + 63) / 64, mx[b_ax] - 1), 64, 0, ctx->stream>>>(L, ctx->geom, ctx->gbc, f, ctx->gbc_off[f]);
    }
    ctx->launches += 6;
  }
  if (nodes_out) F3D_CUDA(cudaMemcpyAsync(nodes_out, nv.p, sizeof(double) * 3 * n_nodes, cudaMemcpyDeviceToHost, ctx->stream));
  F3D_CUDA(cudaStreamSynchronize(ctx->stream));
  F3D_CUDA(cudaGetLastError());
  cudaFree(nv.p); cudaFree(d_grid);
  return 0;
}

int download_records(Ctx* ctx, const double* field0, int n0, int n1, int n2, double* host) {
  const Layout& L = ctx->P.L;
  const size_t n = (size_t)4 * n0 * n1 * n2;
  double* d = nullptr;
  F3D_CUDA(cudaMalloc((void**)&d, n * sizeof(double)));
  k_records_out<<<dim3((n0 + 63) / 64, n1, n2), 64, 0, ctx->stream>>>(L, field0, d, n0, n1, n2);
  ctx->launches++;
  F3D_CUDA(cudaMemcpyAsync(host, d, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  F3D_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d);
  return 0;
}

}  // namespace f3d
