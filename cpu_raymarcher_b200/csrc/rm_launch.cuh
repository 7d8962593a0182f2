// rm_launch.cuh — host-side launcher shared by the two kernel translation units.
// Included with RM_NUM (numeric policy) and RM_SUFFIX (val / fast) defined.
#pragma once
#include <algorithm>
#include <mutex>

#include "rm_device.cuh"

namespace rm {

template <class NP, int ACCEL, int PK, bool TCK = false, int ALGT = -1>
static int launch_one(const RenderParams& p, int n_sms, cudaStream_t stream) {
    auto kern = render_kernel<NP, ACCEL, PK, TCK, ALGT>;
    constexpr int kWarps = CtaShape<NP, ACCEL, PK>::kWarps;
    constexpr int kThreads = 32 * kWarps;
    // per-warp double-buffered TMA stages; the tensor-core instance with the static cooperative queue never streams primitives
    // through them and only needs its A tile + B ring + work-item list
    constexpr size_t kDynSmem = NP::kExact ? 0
                                : (TCK && RM_TC_STATIC_QUEUE) ? (size_t)kTcATileBytes + (size_t)kTcStages * kTcTileBytes + 4u * (size_t)kTcItemCap
                                                              : (size_t)kWarps * 2 * kStageBytes;
    // a scene that stays resident in shared memory (the kernel's own rule) only needs the first stage: one copy per CTA
    constexpr int kPerStage0 = kStageBytes / (16 * ((PK == PK_TSPHERE) ? 1 : 4));
    const bool resident = !NP::kExact && !(TCK && RM_TC_STATIC_QUEUE) && p.scene.n_prims > 0 &&
                          ((PK == PK_TSPHERE) ? p.scene.n_chunks <= 4 : p.scene.n_prims <= kPerStage0);
#ifdef RM_FULL_STAGES  // A/B switch: always the full double-buffered allocation
    const size_t dynSmem = NP::kExact ? (size_t)kWarps * 2 * kStageBytes : ((void)resident, kDynSmem);
#else
    const size_t dynSmem = resident ? (size_t)kStageBytes : kDynSmem;  // resident: one copy of the scene per CTA
#endif
    // per instantiation; function attributes are PER DEVICE, and several devices launch from several host threads (rm_pool)
    static std::mutex mu;
    static int blocksPerSM = 0;
    static unsigned long long devDone = 0ull;
    {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return (int)e;
        std::lock_guard<std::mutex> lk(mu);
        if (!((devDone >> (dev & 63)) & 1ull)) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(kDynSmem, (size_t)kWarps * 2 * kStageBytes));
            if (e != cudaSuccess) return (int)e;
            int b = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, kThreads, kDynSmem);
            if (e != cudaSuccess) return (int)e;
            blocksPerSM = std::max(b, 1);
            // the tensor-core search allocates all 512 TMEM columns of the SM: exactly one (16-warp) CTA per SM
            if (!NP::kExact && ACCEL == RM_ACCEL_BVH && PK == PK_TSPHERE) blocksPerSM = 1;
            devDone |= 1ull << (dev & 63);
        }
    }
    // persistent grid: a multiple of the SM count, never more warps than there are tiles
    long long warpsWanted = p.n_tiles;
    int maxBlocks = n_sms * blocksPerSM;
    int blocks = (int)std::min<long long>(maxBlocks, (warpsWanted + kWarps - 1) / kWarps);
    if (blocks < 1) blocks = 1;
    (void)cudaGetLastError();  // a stale error of an unrelated earlier call on this thread must not be blamed on this launch
    kern<<<blocks, kThreads, dynSmem, stream>>>(p);
    return (int)cudaGetLastError();
}

template <class NP>
static int launch_render_t(const RenderParams& p, int n_sms, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const int ak = p.scene.accel_kind;
    const int pk = NP::kExact ? PK_GENERAL : p.scene.prim_kind;
    if (pk == PK_TSPHERE) {
        if constexpr (!NP::kExact) {
            if (ak == RM_ACCEL_NONE) return launch_one<NP, RM_ACCEL_NONE, PK_TSPHERE>(p, n_sms, stream);
            if (ak == RM_ACCEL_OCTREE) return launch_one<NP, RM_ACCEL_OCTREE, PK_TSPHERE>(p, n_sms, stream);
            // cluster data uploaded (>= 256 spheres, RM_DISABLE_TC unset): the tensor-core instance, specialised for the
            // sphere tracer (every BASELINE config's algorithm); otherwise the FFMA-search instance
            if (p.scene.tc_tiles != nullptr && p.scene.n_prims >= 256) {
                if (p.algorithm == RM_ALG_SPHERE_TRACER) return launch_one<NP, RM_ACCEL_BVH, PK_TSPHERE, true, RM_ALG_SPHERE_TRACER>(p, n_sms, stream);
                return launch_one<NP, RM_ACCEL_BVH, PK_TSPHERE, true, -1>(p, n_sms, stream);
            }
            return launch_one<NP, RM_ACCEL_BVH, PK_TSPHERE>(p, n_sms, stream);
        }
    }
    if (ak == RM_ACCEL_NONE) return launch_one<NP, RM_ACCEL_NONE, PK_GENERAL>(p, n_sms, stream);
    if (ak == RM_ACCEL_OCTREE) return launch_one<NP, RM_ACCEL_OCTREE, PK_GENERAL>(p, n_sms, stream);
    return launch_one<NP, RM_ACCEL_BVH, PK_GENERAL>(p, n_sms, stream);
}

template <class NP>
static int launch_shade_t(const ShadeParams& p, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (p.n_pixels <= 0) return 0;
    int blocks = (p.n_pixels + 255) / 256;
    shade_kernel<NP><<<blocks, 256, 0, stream>>>(p);
    return (int)cudaGetLastError();
}

}  // namespace rm
