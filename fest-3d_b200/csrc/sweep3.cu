// Generation-3 fused sweep: the common instantiations (no pressure-based switching, no transition model) and the dispatcher.
// The kernel template lives in sweep3_kernel.cuh; the instantiations with the rare options compile in sweep3_rare.cu.
#include "sweep3_kernel.cuh"

namespace f3d {

int launch_sweep3_rare(Ctx* ctx, KArgs& a);

int sweep3_grid_ctas(const Layout& L) {
  const int chunk = g3::pick_kchunk(L);
  return ((L.imx - 1 + g3::TX - 1) / g3::TX) * ((L.jmx - 1 + g3::TY - 1) / g3::TY) * ((L.kmx - 1 + chunk - 1) / chunk);
}

int launch_sweep3(Ctx* ctx, KArgs& a) {
  const bool rare = ctx->P.trans_bc || ctx->P.pb_switch[0] || ctx->P.pb_switch[1] || ctx->P.pb_switch[2] || ctx->P.kkl || ctx->P.lctm;
  return rare ? launch_sweep3_rare(ctx, a) : g3::launch_sweep3_set<false>(ctx, a);
}

}  // namespace f3d

#ifdef F3D_PHASE_TIMING
extern "C" void fest3d_gpu_phase_dump3() {
  unsigned long long h[32];
  cudaMemcpyFromSymbol(h, f3d::g3::g_phase3, sizeof(h));
  const char* nm[16] = {"I0", "I1", "I2", "I3", "J0", "J1", "J2", "J3", "K0", "K1", "K2", "K3", "IH", "JH", "JL+cell", "C cell"};
  for (int w = 0; w < 16; ++w) {
    const double tot = (double)(h[2 * w] + h[2 * w + 1]);
    printf("g3 warp %2d %-8s work %5.1f %%  wait at the plane barrier %5.1f %%  (total %.3e cycles)\n", w, nm[w], 100.0 * h[2 * w] / (tot > 0 ? tot : 1), 100.0 * h[2 * w + 1] / (tot > 0 ? tot : 1), tot);
  }
}
#endif
