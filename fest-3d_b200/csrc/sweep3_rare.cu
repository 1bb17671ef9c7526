// Generation-3 fused sweep: the instantiations that carry the code of the rare options (MUSCL / PPM pressure-based switching,
// transition = bc); see sweep3_kernel.cuh.
#ifdef F3D_STAGE_CPASYNC   // the previous staging (per-thread cp.async), kept for A/B measurements: make EXTRA=-DF3D_STAGE_CPASYNC
#include "sweep3_kernel_cpasync.cuh"
#else
#include "sweep3_kernel.cuh"
#endif

namespace f3d {

int launch_sweep3_rare(Ctx* ctx, KArgs& a) { return g3::launch_sweep3_set<true>(ctx, a); }

}  // namespace f3d
