// Probe: tensor memory (TMEM) as warp-quarter-private scratch for FP64 values via tcgen05.st / tcgen05.ld (32x32b shape: every
// thread of a warp reads / writes its own lane, columns = 32-bit words).  Checks (1) values written by warp w are read back by a
// DIFFERENT warp of the same lane quarter ((w + 4) % 16) after a CTA barrier, (2) throughput of st + ld of one double per lane.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_st2(uint32_t addr, double v) {
  const uint32_t lo = __double2loint(v), hi = __double2hiint(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ double tm_ld2(uint32_t addr) {
  uint32_t lo, hi;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ void tm_ld8(uint32_t addr, double (&v)[4]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 4; ++i) v[i] = __hiloint2double(r[2 * i + 1], r[2 * i]);
}

__global__ void __launch_bounds__(512, 1) k_probe(double* out, int* bad, long long* clk, int reps) {
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  if (wid == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = tmem_base;
  const uint32_t lane_base = (uint32_t)(32 * (wid & 3)) << 16;      // lane field: the quarter this warp may touch
  const uint32_t my_cols = (uint32_t)(wid >> 2) * 64;               // 4 warps per quarter share the 512 columns: 128 each (64 doubles)
  // (1) write 16 doubles per thread, barrier, read what the neighbour warp of the same quarter wrote
  for (int c = 0; c < 16; ++c) tm_st2(base + lane_base + my_cols + 2 * c, 1000.0 * wid + 10.0 * c + 0.001 * lane);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const int ow = (wid + 4) & 15;
  int nbad = 0;
  for (int c = 0; c < 16; ++c) {
    const double v = tm_ld2(base + lane_base + (uint32_t)(ow >> 2) * 64 + 2 * c);
    if (v != 1000.0 * ow + 10.0 * c + 0.001 * lane) ++nbad;
    if (blockIdx.x == 0) out[(wid * 16 + c) * 32 + lane] = v;
  }
  double v4[4];
  tm_ld8(base + lane_base + (uint32_t)(ow >> 2) * 64, v4);
  for (int c = 0; c < 4; ++c) if (v4[c] != 1000.0 * ow + 10.0 * c + 0.001 * lane) ++nbad;
  if (nbad) atomicAdd(bad, nbad);
  __syncthreads();
  // (2) throughput: every warp st + ld one double per lane, reps times, dependent chain through the value
  double x = lane;
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    tm_st2(base + lane_base + my_cols + 2 * (r & 15), x);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    x = tm_ld2(base + lane_base + my_cols + 2 * (r & 15)) + 1.0;
  }
  const long long t1 = clock64();
  // independent loads only (8 per group, one wait)
  double acc = 0.0;
  const long long t2 = clock64();
  for (int r = 0; r < reps; ++r) {
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo[c]), "=r"(hi[c]) : "r"(base + lane_base + my_cols + 2 * c) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int c = 0; c < 8; ++c) acc += __hiloint2double(hi[c], lo[c]);
  }
  const long long t3 = clock64();
  if (tid == 0 && blockIdx.x == 0) { clk[0] = t1 - t0; clk[1] = t3 - t2; }
  if (x + acc == -1.0) out[0] = x;
  __syncthreads();
  if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(512));
}

int main() {
  double* out; int* bad; long long* clk;
  cudaMalloc(&out, 16 * 16 * 32 * sizeof(double)); cudaMalloc(&bad, 4); cudaMalloc(&clk, 16);
  cudaMemset(bad, 0, 4);
  const int reps = 2000;
  k_probe<<<148, 512>>>(out, bad, clk, reps);
  cudaError_t e = cudaDeviceSynchronize();
  int hb = -1; long long hc[2] = {0, 0};
  cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost); cudaMemcpy(hc, clk, 16, cudaMemcpyDeviceToHost);
  printf("tmem probe: %s, mismatches %d; st+wait+ld+wait round trip %.1f cycles per iteration (16 warps concurrently); 8 independent ld.x2 + wait: %.1f cycles per group "
         "(= %.2f cycles per warp-double with 16 warps in flight)\n", cudaGetErrorString(e), hb, (double)hc[0] / reps, (double)hc[1] / reps, (double)hc[1] / reps / 8.0);
  return (e != cudaSuccess || hb != 0) ? 1 : 0;
}
