// DFMA issue behaviour on B200: throughput per SM as a function of warps per SMSP and ILP per warp.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, double b, int iters) {
  double x[ILP];
#pragma unroll
  for (int u = 0; u < ILP; ++u) x[u] = 1.0 + threadIdx.x * 1e-3 + u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int u = 0; u < ILP; ++u) x[u] = fma(x[u], b, 1e-9);
  }
  double s = 0; 
#pragma unroll
  for (int u = 0; u < ILP; ++u) s += x[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> void run(int warps_per_sm) {
  double* out; cudaMalloc(&out, 148 * 1024 * sizeof(double));
  const int iters = 4000;
  k<ILP><<<148, warps_per_sm * 32>>>(out, 0.999999, 10);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<ILP><<<148, warps_per_sm * 32>>>(out, 0.999999, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double instr_per_warp = (double)iters * 16 * ILP;
  const double cycles = ms * 1e-3 * 1.965e9;
  printf("warps/SM %2d ILP %d: %.2f cycles per DFMA per warp (dependent spacing %.1f), SM rate %.1f lane-ops/clk\n", warps_per_sm, ILP,
         cycles / instr_per_warp, cycles / (iters * 16.0), instr_per_warp * warps_per_sm * 32 / cycles);
  cudaFree(out);
}
int main() {
  for (int w : {4, 8, 12, 16}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
  return 0;
}
