// Wall distance on the device: find_wall_dist of the reference (src/wall/wall_dist.f90:84-131) -- for every node of the block
// (ghost nodes included, -2:imx+3) the minimum distance to the no-slip surface nodes of ALL blocks (brute force, as the
// reference does it), then the cell value as 0.125 x the sum of its eight node values.  This is SURVEY 8(f) rank 1: at 256^3 with
// four 257^2 walls it is 1.8e7 nodes x 2.6e5 wall nodes = 4.8e12 distance evaluations, the dominant start-up cost of an SST run
// on the host and an embarrassingly parallel min-reduction here.
//
// Kernel shape: a thread keeps NPT nodes (coordinates and running minima) in registers and walks the wall nodes, which a CTA
// stages through shared memory in tiles (every lane reads the same wall node: a broadcast LDS).  The minimum is taken on the
// SQUARED distance -- sqrt is monotonic and correctly rounded, so sqrt(min d2) == min sqrt(d2) bit for bit -- and, because
// squared distances are non-negative, as a 64-bit INTEGER minimum of the bit patterns, which keeps the compare off the FP64
// pipe (3 DADD + 1 DMUL + 2 DFMA per pair remain: that pipe bounds the kernel).
#include "ctx.hpp"

namespace f3d {

constexpr int WD_NPT = 4;        // nodes per thread
constexpr int WD_TILE = 1024;    // wall nodes per shared-memory tile (24 KB)
constexpr int WD_THREADS = 256;

__global__ void __launch_bounds__(WD_THREADS) k_node_wall_dist(const double* __restrict__ nodes, long long n_nodes, const double* __restrict__ wall,
                                                                long long n_wall, double* __restrict__ node_dist) {
  __shared__ double sw[3 * WD_TILE];
  double x[WD_NPT], y[WD_NPT], z[WD_NPT];
  long long best[WD_NPT];
  const long long first = ((long long)blockIdx.x * WD_THREADS + threadIdx.x);
  const long long stride = (long long)gridDim.x * WD_THREADS;
#pragma unroll
  for (int p = 0; p < WD_NPT; ++p) {
    const long long n = first + p * stride;
    const long long m = n < n_nodes ? n : n_nodes - 1;
    x[p] = nodes[3 * m]; y[p] = nodes[3 * m + 1]; z[p] = nodes[3 * m + 2];
    best[p] = __double_as_longlong(1.e+20 * 1.e+20);   // node_dist = 1.e+20 before the loop (wall_dist.f90:98)
  }
  for (long long t0 = 0; t0 < n_wall; t0 += WD_TILE) {
    const int nt = (int)min((long long)WD_TILE, n_wall - t0);
    __syncthreads();
    for (int s = threadIdx.x; s < 3 * nt; s += WD_THREADS) sw[s] = wall[3 * t0 + s];
    __syncthreads();
#pragma unroll 4
    for (int s = 0; s < nt; ++s) {
      const double wx = sw[3 * s], wy = sw[3 * s + 1], wz = sw[3 * s + 2];
#pragma unroll
      for (int p = 0; p < WD_NPT; ++p) {
        const double dx = wx - x[p], dy = wy - y[p], dz = wz - z[p];
        const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
        best[p] = min(best[p], __double_as_longlong(d2));
      }
    }
  }
#pragma unroll
  for (int p = 0; p < WD_NPT; ++p) {
    const long long n = first + p * stride;
    if (n < n_nodes) node_dist[n] = (n_wall > 0) ? sqrt(__longlong_as_double(best[p])) : 1.e+20;
  }
}

// dist(i,j,k) = 0.125*(n(i,j,k) + n(i,j+1,k) + n(i,j+1,k+1) + n(i,j,k+1) + n(i+1,j,k+1) + n(i+1,j,k) + n(i+1,j+1,k) + n(i+1,j+1,k+1))
// in that order (wall_dist.f90:114-128), cells -2..imx+2; written to the context's wall-distance field and, optionally, to a
// contiguous array in the reference layout
__global__ void k_cell_wall_dist(const Layout L, const double* __restrict__ nd, double* __restrict__ field, double* __restrict__ flat) {
  const int i = -2 + blockIdx.x * blockDim.x + threadIdx.x, j = -2 + blockIdx.y * blockDim.y + threadIdx.y, k = -2 + blockIdx.z;
  if (i > L.imx + 2 || j > L.jmx + 2) return;
  const long long ni = L.imx + 6, nj = L.jmx + 6;
  auto at = [&](int a, int b, int c) { return nd[(a + 2) + ni * ((b + 2) + nj * (long long)(c + 2))]; };
  double s = at(i, j, k);
  s = s + at(i, j + 1, k); s = s + at(i, j + 1, k + 1); s = s + at(i, j, k + 1);
  s = s + at(i + 1, j, k + 1); s = s + at(i + 1, j, k); s = s + at(i + 1, j + 1, k); s = s + at(i + 1, j + 1, k + 1);
  const double d = 0.125 * s;
  field[L.idx(i, j, k)] = d;
  if (flat) flat[(i + 2) + (long long)(L.imx + 5) * ((j + 2) + (long long)(L.jmx + 5) * (k + 2))] = d;
}

int launch_wall_distance(Ctx* ctx, const double* nodes_host, const double* wall_host, long long n_wall, double* dist_out, double* kernel_ms) {
  const Layout& L = ctx->P.L;
  const long long n_nodes = (long long)(L.imx + 6) * (L.jmx + 6) * (L.kmx + 6);
  const long long n_cells = (long long)(L.imx + 5) * (L.jmx + 5) * (L.kmx + 5);
  double *d_nodes = nullptr, *d_wall = nullptr, *d_nd = nullptr, *d_flat = nullptr;
  F3D_CUDA(cudaMalloc((void**)&d_nodes, sizeof(double) * 3 * n_nodes));
  F3D_CUDA(cudaMalloc((void**)&d_nd, sizeof(double) * n_nodes));
  F3D_CUDA(cudaMalloc((void**)&d_wall, sizeof(double) * 3 * (n_wall > 0 ? n_wall : 1)));
  if (dist_out) F3D_CUDA(cudaMalloc((void**)&d_flat, sizeof(double) * n_cells));
  F3D_CUDA(cudaMemcpyAsync(d_nodes, nodes_host, sizeof(double) * 3 * n_nodes, cudaMemcpyHostToDevice, ctx->stream));
  if (n_wall > 0) F3D_CUDA(cudaMemcpyAsync(d_wall, wall_host, sizeof(double) * 3 * n_wall, cudaMemcpyHostToDevice, ctx->stream));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const long long per_cta = (long long)WD_THREADS * WD_NPT;
  const int grid = (int)((n_nodes + per_cta - 1) / per_cta);
  cudaEventRecord(e0, ctx->stream);
  k_node_wall_dist<<<grid, WD_THREADS, 0, ctx->stream>>>(d_nodes, n_nodes, d_wall, n_wall, d_nd);
  cudaEventRecord(e1, ctx->stream);
  dim3 b2(32, 4, 1), g2((L.imx + 5 + 31) / 32, (L.jmx + 5 + 3) / 4, L.kmx + 5);
  k_cell_wall_dist<<<g2, b2, 0, ctx->stream>>>(L, d_nd, ctx->geom + (long long)G_DIST * L.fs, d_flat);
  ctx->launches += 2;
  if (dist_out) F3D_CUDA(cudaMemcpyAsync(dist_out, d_flat, sizeof(double) * n_cells, cudaMemcpyDeviceToHost, ctx->stream));
  F3D_CUDA(cudaStreamSynchronize(ctx->stream));
  F3D_CUDA(cudaGetLastError());
  if (kernel_ms) { float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); *kernel_ms = ms; }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d_nodes); cudaFree(d_wall); cudaFree(d_nd); if (d_flat) cudaFree(d_flat);
  return 0;
}

}  // namespace f3d
