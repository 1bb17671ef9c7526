// Generation-4 fused sweep: the common instantiations (no pressure-based switching, no transition model) and the dispatcher.
// The kernel template lives in fused_kernel.cuh; the instantiations with the rare options compile in fused_rare.cu.
#include "fused_kernel.cuh"

namespace f3d {

int launch_fused_rare(Ctx* ctx, KArgs& a);

int fused_grid_ctas(const Layout& L) {
  const int chunk = g4::pick_kchunk(L);
  return ((L.imx - 1 + g4::TX - 1) / g4::TX) * ((L.jmx - 1 + g4::TY - 1) / g4::TY) * ((L.kmx - 1 + chunk - 1) / chunk);
}

int launch_fused(Ctx* ctx, KArgs& a) {
  const bool rare = ctx->P.trans_bc || ctx->P.pb_switch[0] || ctx->P.pb_switch[1] || ctx->P.pb_switch[2];
  return rare ? launch_fused_rare(ctx, a) : g4::launch_fused_set<false>(ctx, a);
}

}  // namespace f3d

#ifdef F3D_PHASE_TIMING
extern "C" void fest3d_gpu_phase_dump() {
  unsigned long long h[64];
  cudaMemcpyFromSymbol(h, f3d::g4::g_phase, sizeof(h));
  const char* nm[4] = {"phase 1 work", "wait barrier A", "phase 2 work", "wait barrier B"};
  for (int w = 0; w < 16; ++w) {
    unsigned long long tot = 0;
    for (int n = 0; n < 4; ++n) tot += h[w * 4 + n];
    printf("warp %2d:", w);
    for (int n = 0; n < 4; ++n) printf("  %s %5.1f %%", nm[n], 100.0 * h[w * 4 + n] / (double)(tot ? tot : 1));
    printf("   (total %.3e cycles)\n", (double)tot);
  }
}
#endif
