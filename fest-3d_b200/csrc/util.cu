// Small streaming kernels around the fused residual kernel: temperature refresh, whole-array RK blends and copies,
// halo pack / unpack, block-local minimum of the time step, ghost-shell carry-over between the two state buffers and
// the one-time AoS -> SoA conversion of the reference's geometry records.
//
// Reference: src/update.f90:170 (Temp), :177-178 (U_store = qp, R_store = 0), :203,206,214 (TVD-RK whole-array blends);
// src/interface1.f90:96-493 (pack order n -> layer -> outer transverse -> inner transverse; unpack with PxDir / dir_switch);
// src/time.f90:286 (delta_t = minval(delta_t), block-local).
#include "ctx.hpp"

namespace f3d {

__global__ void k_temp(const Params P, const double* __restrict__ q, double* __restrict__ temp) {
  const Layout& L = P.L;
  const int i = -2 + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = -2 + blockIdx.y * blockDim.y + threadIdx.y;
  const int k = -2 + blockIdx.z;
  if (i > L.imx + 2 || j > L.jmx + 2) return;
  const long long c = L.idx(i, j, k);
  temp[c] = q[4 * L.fs + c] / (P.R_gas * q[c]);
}

int launch_temp(Ctx* ctx) {
  const Layout& L = ctx->P.L;
  dim3 block(64, 4), grid((L.imx + 5 + 63) / 64, (L.jmx + 5 + 3) / 4, L.kmx + 5);
  k_temp<<<grid, block, 0, ctx->stream>>>(ctx->P, ctx->qp, ctx->temp);
  ctx->launches++;
  F3D_CUDA(cudaGetLastError());
  return 0;
}

// whole padded fields, flat
__global__ void k_blend(double* __restrict__ q, const double* __restrict__ u, double a, double b, long long n) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) q[t] = a * u[t] + b * q[t];
}
__global__ void k_copy(double* __restrict__ d, const double* __restrict__ s, long long n) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) d[t] = s[t];
}
__global__ void k_zero(double* __restrict__ d, long long n) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) d[t] = 0.0;
}

static inline int flat_grid(long long n) { long long g = (n + 255) / 256; return (int)(g > 148 * 16 ? 148 * 16 : g); }

int launch_blend(Ctx* ctx, double a, double b) {
  const long long n = ctx->P.L.fs * ctx->P.L.nv;
  k_blend<<<flat_grid(n), 256, 0, ctx->stream>>>(ctx->qp, ctx->ustore, a, b, n);
  ctx->launches++;
  F3D_CUDA(cudaGetLastError());
  return 0;
}
int launch_copy_fields(Ctx* ctx, double* dst, const double* src, int nfields) {
  const long long n = ctx->P.L.fs * nfields;
  k_copy<<<flat_grid(n), 256, 0, ctx->stream>>>(dst, src, n);
  ctx->launches++;
  F3D_CUDA(cudaGetLastError());
  return 0;
}
int launch_zero_fields(Ctx* ctx, double* dst, int nfields) {
  const long long n = ctx->P.L.fs * nfields;
  k_zero<<<flat_grid(n), 256, 0, ctx->stream>>>(dst, n);
  ctx->launches++;
  F3D_CUDA(cudaGetLastError());
  return 0;
}

// every cell that is not interior: dst <- src (so that the new-state buffer is the complete array the reference
// would hold after its in-place update).  Only the shell is enumerated (a full-volume launch with an early exit for
// the interior cost 83 us per stage at 256^3): launch 1 copies the six ghost planes whole, launch 2 the frame of every
// interior plane (six ghost rows over the whole i extent, six ghost columns of the interior rows).
__global__ void k_ghost_planes(const Params P, double* __restrict__ dst, const double* __restrict__ src) {
  const Layout& L = P.L;
  const int ni = L.imx + 5, nj = L.jmx + 5;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ni * nj) return;
  const int k = (blockIdx.y < 3) ? -2 + (int)blockIdx.y : L.kmx + ((int)blockIdx.y - 3);
  const long long c = L.idx(-2 + t % ni, -2 + t / ni, k);
  for (int v = 0; v < L.nv; ++v) dst[v * L.fs + c] = src[v * L.fs + c];
}
__global__ void k_ghost_frames(const Params P, double* __restrict__ dst, const double* __restrict__ src) {
  const Layout& L = P.L;
  const int ni = L.imx + 5;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 6 * ni + 6 * (L.jmx - 1)) return;
  const int k = 1 + blockIdx.y;
  int i, j;
  if (t < 6 * ni) { const int r = t / ni; i = -2 + t % ni; j = (r < 3) ? -2 + r : L.jmx + (r - 3); }
  else { const int u = t - 6 * ni, cidx = u % 6; j = 1 + u / 6; i = (cidx < 3) ? -2 + cidx : L.imx + (cidx - 3); }
  const long long c = L.idx(i, j, k);
  for (int v = 0; v < L.nv; ++v) dst[v * L.fs + c] = src[v * L.fs + c];
}

int launch_ghost_shell_copy(Ctx* ctx, double* dst, const double* src) {
  const Layout& L = ctx->P.L;
  const int plane = (L.imx + 5) * (L.jmx + 5), frame = 6 * (L.imx + 5) + 6 * (L.jmx - 1);
  k_ghost_planes<<<dim3((plane + 255) / 256, 6), 256, 0, ctx->stream>>>(ctx->P, dst, src);
  k_ghost_frames<<<dim3((frame + 255) / 256, L.kmx - 1), 256, 0, ctx->stream>>>(ctx->P, dst, src);
  ctx->launches += 2;
  F3D_CUDA(cudaGetLastError());
  return 0;
}

// ---- halo pack / unpack ------------------------------------------------------------------------------------------
// buffer order (interface1.f90:130-138): n (variable) slowest, then layer l = 1..3, then the outer transverse axis,
// then the inner transverse axis.  Faces 1,2: inner j, outer k; 3,4: inner i, outer k; 5,6: inner i, outer j.
__global__ void k_pack(const Params P, const double* __restrict__ q, double* __restrict__ buf, int face) {
  const Layout& L = P.L;
  const int ax = (face - 1) / 2;
  const bool lo = (face % 2) == 1;
  const int a_ax = (ax == 0) ? 1 : 0, b_ax = (ax == 2) ? 1 : 2;
  const int mx[3] = {L.imx, L.jmx, L.kmx};
  const int na = mx[a_ax] - 1, nb = mx[b_ax] - 1;
  const long long per = (long long)na * nb;
  const long long total = per * 3 * L.nv;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(t % na), b = (int)((t / na) % nb), l = (int)((t / per) % 3) + 1, n = (int)(t / (per * 3));
    int idx[3]; idx[a_ax] = a + 1; idx[b_ax] = b + 1; idx[ax] = lo ? l : mx[ax] - l;
    buf[t] = q[n * L.fs + L.idx(idx[0], idx[1], idx[2])];
  }
}

struct UnpackMap { int alo, ahi, adir, blo, bhi, bdir, dir_switch; };

__global__ void k_unpack(const Params P, double* __restrict__ q, const double* __restrict__ buf, int face, UnpackMap m) {
  const Layout& L = P.L;
  const int ax = (face - 1) / 2;
  const bool lo = (face % 2) == 1;
  const int a_ax = (ax == 0) ? 1 : 0, b_ax = (ax == 2) ? 1 : 2;
  const int mx[3] = {L.imx, L.jmx, L.kmx};
  const int na = (m.adir > 0 ? m.ahi - m.alo : m.alo - m.ahi) + 1;
  const int nb = (m.bdir > 0 ? m.bhi - m.blo : m.blo - m.bhi) + 1;
  const long long per = (long long)na * nb;
  const long long total = per * 3 * L.nv;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int l = (int)((t / per) % 3) + 1, n = (int)(t / (per * 3));
    const long long r = t % per;
    int ia, ib;   // loop counters along a and b in the order the reference's unpack loops run
    if (m.dir_switch == 0) { ia = (int)(r % na); ib = (int)(r / na); }   // outer b, inner a
    else { ib = (int)(r % nb); ia = (int)(r / nb); }                     // outer a, inner b
    int idx[3];
    idx[a_ax] = m.alo + ia * m.adir; idx[b_ax] = m.blo + ib * m.bdir; idx[ax] = lo ? 1 - l : mx[ax] + l - 1;
    q[n * L.fs + L.idx(idx[0], idx[1], idx[2])] = buf[t];
  }
}

int launch_pack(Ctx* ctx, int face) {
  const long long total = (long long)ctx->buf_elems[face - 1];
  k_pack<<<flat_grid(total), 256, 0, ctx->stream>>>(ctx->P, ctx->qp, ctx->sendbuf[face - 1], face);
  ctx->launches++;
  F3D_CUDA(cudaGetLastError());
  return 0;
}

int launch_unpack(Ctx* ctx, int face, const double* buf) {
  const int f = face - 1;
  UnpackMap m{ctx->cfg.plo[f][0], ctx->cfg.phi[f][0], ctx->cfg.pdir[f][0], ctx->cfg.plo[f][1], ctx->cfg.phi[f][1], ctx->cfg.pdir[f][1], ctx->cfg.dir_switch[f]};
  const long long total = (long long)ctx->buf_elems[f];
  k_unpack<<<flat_grid(total), 256, 0, ctx->stream>>>(ctx->P, ctx->qp, buf, face, m);
  ctx->launches++;
  F3D_CUDA(cudaGetLastError());
  return 0;
}

// ---- block-local min of delta_t over interior cells, then broadcast (time.f90:286) -------------------------------
__global__ void k_min_partial(const Params P, const double* __restrict__ dt, double* __restrict__ part) {
  const Layout& L = P.L;
  const long long n = (long long)(L.imx - 1) * (L.jmx - 1) * (L.kmx - 1);
  double m = 1e300;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t % (L.imx - 1)) + 1, j = (int)((t / (L.imx - 1)) % (L.jmx - 1)) + 1, k = (int)(t / ((long long)(L.imx - 1) * (L.jmx - 1))) + 1;
    m = fmin(m, dt[L.idx(i, j, k)]);
  }
  __shared__ double sm[8];
  for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_down_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmin(m, sm[w]);
    part[blockIdx.x] = m;
  }
}
__global__ void k_min_fill(const Params P, double* __restrict__ dt, const double* __restrict__ part, int nparts) {
  const Layout& L = P.L;
  double m = 1e300;
  for (int p = 0; p < nparts; ++p) m = fmin(m, part[p]);
  const long long n = (long long)(L.imx - 1) * (L.jmx - 1) * (L.kmx - 1);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t % (L.imx - 1)) + 1, j = (int)((t / (L.imx - 1)) % (L.jmx - 1)) + 1, k = (int)(t / ((long long)(L.imx - 1) * (L.jmx - 1))) + 1;
    dt[L.idx(i, j, k)] = m;
  }
}

int launch_global_dt(Ctx* ctx) {
  const int nparts = 128;
  k_min_partial<<<nparts, 256, 0, ctx->stream>>>(ctx->P, ctx->dt, ctx->red);
  k_min_fill<<<296, 256, 0, ctx->stream>>>(ctx->P, ctx->dt, ctx->red, nparts);
  ctx->launches += 2;
  F3D_CUDA(cudaGetLastError());
  return 0;
}

// reference layout qp(-2:imx+2,-2:jmx+2,-2:kmx+2,nv) (contiguous) <-> padded SoA fields
__global__ void k_state_relayout(const Params P, double* __restrict__ fields, double* __restrict__ flat, int to_fields) {
  const Layout& L = P.L;
  const int n0 = L.imx + 5, n1 = L.jmx + 5, n2 = L.kmx + 5;
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y, k = blockIdx.z;
  if (i >= n0 || j >= n1) return;
  const long long c = L.idx(i - 2, j - 2, k - 2);
  const size_t per = (size_t)n0 * n1 * n2, t = (size_t)i + (size_t)n0 * ((size_t)j + (size_t)n1 * k);
  for (int v = 0; v < L.nv; ++v) {
    if (to_fields) fields[v * L.fs + c] = flat[v * per + t];
    else flat[v * per + t] = fields[v * L.fs + c];
  }
}

int launch_state_relayout(Ctx* ctx, double* fields, double* flat, int to_fields) {
  const Layout& L = ctx->P.L;
  dim3 block(64, 4), grid((L.imx + 5 + 63) / 64, (L.jmx + 5 + 3) / 4, L.kmx + 5);
  k_state_relayout<<<grid, block, 0, ctx->stream>>>(ctx->P, fields, flat, to_fields);
  ctx->launches++;
  F3D_CUDA(cudaGetLastError());
  return 0;
}

// ---- one-time layout conversions ----------------------------------------------------------------------------------
// AoS records (4 doubles, reference extents n0 x n1 x n2, lower bound -2) -> four SoA padded fields
__global__ void k_rec_to_fields(const Params P, const double* __restrict__ rec, int n0, int n1, int n2, double* __restrict__ f0, long long fstride) {
  const Layout& L = P.L;
  const long long n = (long long)n0 * n1 * n2;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t % n0) - 2, j = (int)((t / n0) % n1) - 2, k = (int)(t / ((long long)n0 * n1)) - 2;
    const long long c = L.idx(i, j, k);
#pragma unroll
    for (int m = 0; m < 4; ++m) f0[m * fstride + c] = rec[4 * t + m];
  }
}

int upload_records(Ctx* ctx, const double* host, int n0, int n1, int n2, double* field0) {
  const size_t bytes = (size_t)4 * n0 * n1 * n2 * sizeof(double);
  F3D_CUDA(cudaMemcpyAsync(ctx->staging, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  const long long n = (long long)n0 * n1 * n2;
  k_rec_to_fields<<<flat_grid(n), 256, 0, ctx->stream>>>(ctx->P, ctx->staging, n0, n1, n2, field0, ctx->P.L.fs);
  ctx->launches++;
  F3D_CUDA(cudaGetLastError());
  F3D_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // namespace f3d
