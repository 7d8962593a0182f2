// rm_kernels_fast.cu — fp32 fast-path build of the render kernels (FMA contraction on).
#include "rm_launch.cuh"

namespace rm {
int launch_render_fast(const RenderParams& p, int n_sms, void* stream) { return launch_render_t<NumFast>(p, n_sms, stream); }
int launch_shade_fast(const ShadeParams& p, void* stream) { return launch_shade_t<NumFast>(p, stream); }
}  // namespace rm
