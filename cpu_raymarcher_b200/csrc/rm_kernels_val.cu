// rm_kernels_val.cu — fp64 validation build of the render kernels (compiled with -fmad=false).
// JS Number is an IEEE double and is never fused; this TU reproduces that bit for bit.
#include "rm_launch.cuh"

namespace rm {
int launch_render_val(const RenderParams& p, int n_sms, void* stream) { return launch_render_t<NumJS>(p, n_sms, stream); }
int launch_shade_val(const ShadeParams& p, void* stream) { return launch_shade_t<NumJS>(p, stream); }
}  // namespace rm
