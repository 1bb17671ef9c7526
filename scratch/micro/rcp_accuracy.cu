// Accuracy of the fast FP64 reciprocal / rsqrt / sqrt of physics.cuh against the IEEE operations, over 2^24 operands
// spread over 40 binades: prints the maximum relative error in units of 2^-53.
#include <cstdio>
#include <cmath>
#include "../../fest-3d_b200/csrc/physics.cuh"
__global__ void k(double* out, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  double m0 = 0, m1 = 0, m2 = 0;
  for (int s = t; s < n; s += gridDim.x * blockDim.x) {
    const double x = ldexp(1.0 + (double)s * (1.0 / n) + 1e-9 * s, (s % 40) - 20);
    m0 = fmax(m0, fabs(f3d::rcp64(x) * x - 1.0));                       // vs exact 1/x (x*(1/x) - 1 is exact to 2^-53 here)
    m1 = fmax(m1, fabs(f3d::rsqrt64(x) / (1.0 / sqrt(x)) - 1.0));
    m2 = fmax(m2, fabs(f3d::sqrt64(x) / sqrt(x) - 1.0));
  }
  atomicMax((unsigned long long*)&out[0], (unsigned long long)__double_as_longlong(m0));
  atomicMax((unsigned long long*)&out[1], (unsigned long long)__double_as_longlong(m1));
  atomicMax((unsigned long long*)&out[2], (unsigned long long)__double_as_longlong(m2));
}
int main() {
  double* d; cudaMalloc(&d, 24); cudaMemset(d, 0, 24);
  k<<<592, 256>>>(d, 1 << 24);
  double h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  const double u = ldexp(1.0, -53);
  printf("max relative error / 2^-53: rcp64 %.2f  rsqrt64 %.2f  sqrt64 %.2f\n", h[0] / u, h[1] / u, h[2] / u);
  return (h[0] < 8 * u && h[1] < 8 * u && h[2] < 8 * u) ? 0 : 1;
}
