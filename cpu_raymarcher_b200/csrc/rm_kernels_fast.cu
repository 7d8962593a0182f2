// rm_kernels_fast.cu — fp32 fast-path build of the render kernels (FMA contraction on).
#include "rm_launch.cuh"

namespace rm {
int launch_render_fast(const RenderParams& p, int n_sms, void* stream) { return launch_render_t<NumFast>(p, n_sms, stream); }
int launch_shade_fast(const ShadeParams& p, void* stream) { return launch_shade_t<NumFast>(p, stream); }
int launch_order_tiles(const unsigned int* cost, unsigned int* order, int n_tiles, int n_runs, void* stream) {
    if (n_tiles <= 0) return 0;
    n_runs = n_runs < 0 ? 0 : (n_runs > kOrderBands ? kOrderBands : n_runs);
    order_tiles_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(cost, order, n_tiles, n_runs);
    return (int)cudaGetLastError();
}
}  // namespace rm
