// Generation-4 fused sweep: the instantiations that carry the code of the rare options (MUSCL / PPM pressure-based switching,
// transition = bc); see fused_kernel.cuh.
#include "fused_kernel.cuh"

namespace f3d {

int launch_fused_rare(Ctx* ctx, KArgs& a) { return g4::launch_fused_set<true>(ctx, a); }

}  // namespace f3d
