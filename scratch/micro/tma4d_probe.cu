// Probe: one cp.async.bulk.tensor.4d (UTMALDG) per staged array instead of one bulk copy per field row.  A padded FP64 field set
// [field][k][j][i] (pitches sj, sk, fs like ctx.hpp:Layout) is described by a 4-D tensor map; a box (36 x rows x 1 x nfields)
// lands in shared memory as [field][row][col] -- the staged-plane layout of sweep3 -- and is checked element by element.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

constexpr int BX = 36, BY = 6, NFB = 18;   // box: columns, rows, fields

__global__ void probe(const __grid_constant__ CUtensorMap tm, const double* __restrict__ src, long long sj, long long sk, long long fs, int x0, int y0,
                      int nz, int* bad) {
  extern __shared__ __align__(128) double sm[];
  __shared__ __align__(8) unsigned long long mbar[2];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[b])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  constexpr unsigned BYTES = BX * BY * NFB * 8;
  auto issue = [&](int z) {
    if (tid == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar[z & 1])), "r"(BYTES) : "memory");
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                       smem_u32(sm + (z & 1) * BX * BY * NFB)),
                   "l"(&tm), "r"(x0), "r"(y0), "r"(z), "r"(0), "r"(smem_u32(&mbar[z & 1]))
                   : "memory");
    }
  };
  issue(0);
  for (int z = 0; z < nz; ++z) {
    __syncthreads();
    if (z + 1 < nz) issue(z + 1);
    const unsigned par = (z >> 1) & 1;
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(&mbar[z & 1])),
        "r"(par)
        : "memory");
    const double* pl = sm + (z & 1) * BX * BY * NFB;
    for (int s = tid; s < BX * BY * NFB; s += blockDim.x) {
      const int f = s / (BX * BY), r = (s / BX) % BY, c = s % BX;
      const double want = src[f * fs + z * sk + (y0 + r) * sj + x0 + c];
      if (pl[s] != want) atomicAdd(bad, 1);
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const long long sj = 288, pj = 20, pk = 12, sk = sj * pj, fs = ((13 + sk * pk + 31) / 32) * 32;
  const int nf = 18;
  const long long n = fs * nf;
  double* h = (double*)malloc(n * 8);
  for (long long i = 0; i < n; ++i) h[i] = (double)(i % 1000003) * 0.25 + 1.0;
  double* d; int* bad;
  cudaMalloc(&d, n * 8); cudaMalloc(&bad, 4);
  cudaMemcpy(d, h, n * 8, cudaMemcpyHostToDevice); cudaMemset(bad, 0, 4);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (ge != cudaSuccess || !fn) { printf("no cuTensorMapEncodeTiled: %s\n", cudaGetErrorString(ge)); return 2; }
  CUtensorMap tm;
  const cuuint64_t dims[4] = {(cuuint64_t)sj, (cuuint64_t)pj, (cuuint64_t)pk, (cuuint64_t)nf};
  const cuuint64_t strides[3] = {(cuuint64_t)sj * 8, (cuuint64_t)sk * 8, (cuuint64_t)fs * 8};
  const cuuint32_t box[4] = {BX, BY, 1, NFB};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d (fs %% sk = %lld)\n", (int)r, fs % sk);
  if (r != CUDA_SUCCESS) return 3;
  const size_t shm = 2 * BX * BY * NFB * 8;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
  probe<<<8, 256, shm>>>(tm, d, sj, sk, fs, 46, 3, (int)pk, bad);   // x0 = 14 + 32 like idx(i0-2, ...) of the second tile
  cudaError_t e = cudaDeviceSynchronize();
  int hb = -1; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
  printf("tma 4d probe: %s, mismatches %d\n", cudaGetErrorString(e), hb);
  return (e == cudaSuccess && hb == 0) ? 0 : 1;
}
