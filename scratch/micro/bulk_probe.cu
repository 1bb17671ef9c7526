// Probe of the staging mechanism generation 3 is about to adopt: rows of a padded FP64 field copied into shared memory by
// cp.async.bulk (the TMA engine, UBLKCP in SASS), completion on an mbarrier with expect_tx, two buffers reused over many
// phases, one producer warp with the copies spread over its lanes, all other threads as consumers.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(b))
               : "memory");
}

constexpr int ROWS = 40, COLS = 36, NPL = 37;   // rows per plane, doubles per row, planes

__global__ void probe(const double* __restrict__ src, long long pitch, long long plane_stride, double* out, int* bad) {
  extern __shared__ __align__(128) double sm[];
  __shared__ __align__(8) unsigned long long mbar[2];
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  auto issue = [&](int p) {   // producer warp: plane p into buffer p & 1
    double* dst = sm + (p & 1) * ROWS * COLS;
    if (lane == 0) mbar_expect_tx(&mbar[p & 1], ROWS * COLS * 8);
    __syncwarp();
    for (int r = lane; r < ROWS; r += 32) bulk_g2s(dst + r * COLS, src + p * plane_stride + r * pitch, COLS * 8, &mbar[p & 1]);
  };
  if (wid == 0) issue(0);
  double acc = 0.0;
  for (int p = 0; p < NPL; ++p) {
    __syncthreads();                       // everyone is done with plane p-1 (buffer (p+1) & 1)
    if (wid == 0 && p + 1 < NPL) issue(p + 1);
    mbar_wait(&mbar[p & 1], (p >> 1) & 1);
    const double* pl = sm + (p & 1) * ROWS * COLS;
    for (int s = tid; s < ROWS * COLS; s += blockDim.x) {
      const double want = src[p * plane_stride + (s / COLS) * pitch + (s % COLS)];
      if (pl[s] != want) atomicAdd(bad, 1);
      acc += pl[s];
    }
  }
  out[blockIdx.x * blockDim.x + tid] = acc;
}

int main() {
  const long long pitch = 272, plane_stride = pitch * 48 + 32;
  const long long n = plane_stride * (NPL + 1) + 64;
  double* h = (double*)malloc(n * 8);
  for (long long i = 0; i < n; ++i) h[i] = (double)(i % 100003) * 0.5 + 1.0;
  double *d, *out; int* bad;
  cudaMalloc(&d, n * 8); cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&bad, 4);
  cudaMemcpy(d, h, n * 8, cudaMemcpyHostToDevice); cudaMemset(bad, 0, 4);
  const size_t shm = 2 * ROWS * COLS * 8;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
  probe<<<148, 512, shm>>>(d + 14, pitch, plane_stride, out, bad);   // + 14 doubles: an even element offset like idx(i0-2, j, k)
  cudaError_t e = cudaDeviceSynchronize();
  int hb = -1; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
  printf("bulk probe: %s, mismatches %d\n", cudaGetErrorString(e), hb);
  return (e == cudaSuccess && hb == 0) ? 0 : 1;
}
