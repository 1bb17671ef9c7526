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
  const bool rare = ctx->P.trans_bc || ctx->P.pb_switch[0] || ctx->P.pb_switch[1] || ctx->P.pb_switch[2];
  return rare ? launch_sweep3_rare(ctx, a) : g3::launch_sweep3_set<false>(ctx, a);
}

}  // namespace f3d
