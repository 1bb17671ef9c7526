// Generation-3 fused sweep: the instantiations that carry the code of the rare options (MUSCL / PPM pressure-based switching,
// transition = bc); see sweep3_kernel.cuh.
#include "sweep3_kernel.cuh"

namespace f3d {

int launch_sweep3_rare(Ctx* ctx, KArgs& a) { return g3::launch_sweep3_set<true>(ctx, a); }

}  // namespace f3d
