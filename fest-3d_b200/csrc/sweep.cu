// Launch side of the fused residual / time-step / update sweep: kernel arguments, CUDA-event timing of the sweep launch,
// per-CTA norm partials -> Res_abs (resnorm.f90:171-199).  The kernels live in sweep3_kernel.cuh (staged form, the default:
// sweep3.cu / sweep3_rare.cu) and fused_kernel.cuh (F3D_GRADIENTS=fused: fused.cu / fused_rare.cu).
#include "sweep_common.cuh"
#include <cstdlib>
#include <cstring>

namespace f3d {


// final reduction of the per-CTA partials in a fixed order (deterministic), scaled like get_absolute_resnorm.  The iteration slot the
// norms go to is a device-side counter (advanced here), so that the launch is the same every iteration and can sit in a CUDA graph.
__global__ void k_norm_final(const double* __restrict__ red, int n_cta, int nvp1, const double* scale /* nvp1 */, double* norms, int* slot_ctr) {
  __shared__ double sm[32];
  const int slot = *slot_ctr;
  for (int v = 0; v < nvp1; ++v) {
    double x = 0.0;
    for (int b = threadIdx.x; b < n_cta; b += blockDim.x) x += red[(long long)b * nvp1 + v];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
      norms[(long long)slot * nvp1 + v] = (v == 0) ? (t / scale[0]) : (t / (scale[v] * scale[v]));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *slot_ctr = slot + 1;
}

int fused_grid_ctas(const Layout& L);
int launch_fused(Ctx* ctx, KArgs& a);
int sweep3_grid_ctas(const Layout& L);
int launch_sweep3(Ctx* ctx, KArgs& a);

int residual_grid_ctas(const Layout& L) { const int a_ = fused_grid_ctas(L), b_ = sweep3_grid_ctas(L); return a_ > b_ ? a_ : b_; }   // one norm partial per CTA of the sweep

int launch_residual(Ctx* ctx, int mode, double TF, double SF, int use_store_sum, int first_stage, int want_norms) {
  KArgs a{};
  a.q = ctx->qp;
  const bool have_store = ctx->cfg.time_accuracy == F3D_T_RK2 || ctx->cfg.time_accuracy == F3D_T_RK4;
  a.quse = (mode == MODE_UPDATE && have_store) ? ctx->ustore : ctx->qp;
  a.qnew = ctx->qp2;
  a.residue = ctx->residue;
  a.rstore = (mode == MODE_UPDATE && have_store) ? ctx->rstore : nullptr;
  a.dt = ctx->dt;
  a.geom = ctx->geom;
  a.grad = ctx->grad;
  a.mu = ctx->mu;
  a.src = ctx->src;
  a.gbc = ctx->gbc;
  for (int f = 0; f < 6; ++f) a.gbc_off[f] = ctx->gbc_off[f];
  a.red = ctx->red;
  a.err = ctx->err_dev;
  a.mode = mode; a.first_stage = first_stage; a.want_norms = want_norms;
  a.have_store = (mode == MODE_UPDATE && have_store) ? 1 : 0;
  a.use_store_sum = use_store_sum;
  a.TF = TF; a.SF = SF;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->timing) {
    if (ctx->ev_used == ctx->ev_pool.size()) {
      cudaEvent_t x, y; cudaEventCreate(&x); cudaEventCreate(&y);
      ctx->ev_pool.emplace_back(x, y);
    }
    e0 = ctx->ev_pool[ctx->ev_used].first; e1 = ctx->ev_pool[ctx->ev_used].second; ctx->ev_used++;
    cudaEventRecord(e0, ctx->stream);
  }
  const int rc = ctx->fused ? launch_fused(ctx, a) : launch_sweep3(ctx, a);
  if (ctx->timing) cudaEventRecord(e1, ctx->stream);
  if (rc) return rc;
  F3D_CUDA(cudaGetLastError());
  return 0;
}


int launch_norms(Ctx* ctx) {
  const int nvp1 = ctx->P.L.nv + 1;
  k_norm_final<<<1, 256, 0, ctx->stream>>>(ctx->red, ctx->red_blocks, nvp1, ctx->norms_dev + 1024, ctx->norms_dev, ctx->err_dev + 4);
  ctx->launches++;
  F3D_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace f3d
