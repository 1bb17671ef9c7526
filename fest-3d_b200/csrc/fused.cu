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
