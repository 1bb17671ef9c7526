// FP64 op-cost microbenchmark for B200 (sm_100a): lane-ops per clock per SM for the primitives the path leans on.
#include <cstdio>
#include <cuda_runtime.h>
#include <cmath>

__device__ __forceinline__ double rcp_fast(double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0); r = fma(r, e, r);
  e = fma(-b, r, 1.0); r = fma(r, e, r);
  return r;
}
__device__ __forceinline__ double div_fast(double a, double b) {
  const double r = rcp_fast(b);
  double q = a * r;
  return fma(fma(-b, q, a), r, q);
}
__device__ __forceinline__ double rsqrt_fast(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  // two Newton steps: r = r*(1.5 - 0.5*x*r*r)
  double h = 0.5 * x;
  double t = fma(-h * r, r, 0.5); r = fma(r, t, r);
  t = fma(-h * r, r, 0.5); r = fma(r, t, r);
  return r;
}
__device__ __forceinline__ double sqrt_fast(double x) {
  const double r = rsqrt_fast(x);
  double s = x * r;
  return fma(fma(-s, s, x), 0.5 * r, s);
}

template <int OP>
__global__ void __launch_bounds__(256) k(double* out, double a0, double b0, int iters) {
  double x0 = a0 + threadIdx.x * 1e-3, x1 = x0 + 0.1, x2 = x0 + 0.2, x3 = x0 + 0.3;
  const double b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (OP == 0) { x0 = fma(x0, b, 1e-9); x1 = fma(x1, b, 1e-9); x2 = fma(x2, b, 1e-9); x3 = fma(x3, b, 1e-9); }
      if (OP == 1) { x0 = 1.0 + b / x0; x1 = 1.0 + b / x1; x2 = 1.0 + b / x2; x3 = 1.0 + b / x3; }
      if (OP == 2) { x0 = 1.0 + div_fast(b, x0); x1 = 1.0 + div_fast(b, x1); x2 = 1.0 + div_fast(b, x2); x3 = 1.0 + div_fast(b, x3); }
      if (OP == 3) { x0 = 1.0 + sqrt(x0); x1 = 1.0 + sqrt(x1); x2 = 1.0 + sqrt(x2); x3 = 1.0 + sqrt(x3); }
      if (OP == 4) { x0 = 1.0 + sqrt_fast(x0); x1 = 1.0 + sqrt_fast(x1); x2 = 1.0 + sqrt_fast(x2); x3 = 1.0 + sqrt_fast(x3); }
      if (OP == 5) { x0 = fmax(0.5, fmin(x0 * b, 2.0)); x1 = fmax(0.5, fmin(x1 * b, 2.0)); x2 = fmax(0.5, fmin(x2 * b, 2.0)); x3 = fmax(0.5, fmin(x3 * b, 2.0)); }
      if (OP == 6) { x0 = 1.0 + pow(x0, 1.5) * 1e-3; x1 = 1.0 + pow(x1, 1.5) * 1e-3; x2 = 1.0 + pow(x2, 1.5) * 1e-3; x3 = 1.0 + pow(x3, 1.5) * 1e-3; }
      if (OP == 7) { x0 = 1.0 + tanh(x0); x1 = 1.0 + tanh(x1); x2 = 1.0 + tanh(x2); x3 = 1.0 + tanh(x3); }
      if (OP == 8) { x0 = x0 + b; x1 = x1 + b; x2 = x2 + b; x3 = x3 + b; }
      if (OP == 9) { x0 = 1.0 + b * rcp_fast(x0); x1 = 1.0 + b * rcp_fast(x1); x2 = 1.0 + b * rcp_fast(x2); x3 = 1.0 + b * rcp_fast(x3); }
      if (OP == 10) { x0 = copysign(1e-14, x0) + x0 * b; x1 = copysign(1e-14, x1) + x1 * b; x2 = copysign(1e-14, x2) + x2 * b; x3 = copysign(1e-14, x3) + x3 * b; }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}

template <int OP>
void run(const char* name, double a0, double b0, int extra_ops) {
  double* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(double));
  const int iters = 2000, blocks = 148 * 8;
  k<OP><<<blocks, 256>>>(out, a0, b0, 10);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP><<<blocks, 256>>>(out, a0, b0, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)blocks * 256 * iters * 8 * 4;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-28s %8.3f ms  %8.2f Gop/s  %6.2f lane-ops/clk/SM @%d MHz (nominal)  -> %5.1f DFMA-slots/op\n", name, ms, ops / ms / 1e6,
         ops / (ms * 1e-3) / 148.0 / (clk * 1e3), clk / 1000, 64.0 / (ops / (ms * 1e-3) / 148.0 / (clk * 1e3)));
  cudaFree(out);
}

int main() {
  run<0>("dfma", 1.0, 0.999999, 0);
  run<8>("dadd", 1.0, 1e-9, 0);
  run<1>("1+b/x (IEEE div)", 1.5, 0.7, 0);
  run<2>("1+div_fast(b,x)", 1.5, 0.7, 0);
  run<9>("1+b*rcp_fast(x)", 1.5, 0.7, 0);
  run<3>("1+sqrt(x)", 1.5, 0.7, 0);
  run<4>("1+sqrt_fast(x)", 1.5, 0.7, 0);
  run<5>("fmax(.5,fmin(x*b,2))", 1.5, 0.999, 0);
  run<10>("copysign+fma", 1.5, 0.999, 0);
  run<6>("1+pow(x,1.5)*1e-3", 1.5, 0.7, 0);
  run<7>("1+tanh(x)", 0.5, 0.7, 0);
  return 0;
}
