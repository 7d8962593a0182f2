// rm_ctx.h — internal: the context object behind the opaque rm_ctx of include/rm.h (shared by rm_api.cu and rm_pool.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "rm_host.h"
#include "rm_types.h"

namespace rm {
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};
// One device allocation of the resident scene: where it lives, how big it is, and which pointer field of DevScene names it
// (byte offset inside DevScene) — enough to replicate the scene on another device with peer copies (rm_pool_upload_scene).
struct SceneAlloc {
    void* p;
    size_t bytes;
    ptrdiff_t field_off;
};
}  // namespace rm

struct rm_ctx {
    int device = 0;
    unsigned flags = 0;
    int n_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::mutex mu;
    std::string err;
    // scene
    bool has_scene = false;
    bool exact_only = false;  // operator trees / mandelbulb: always the exact (fp64) kernels
    rm::DevScene scene{};
    std::vector<rm::SceneAlloc> scene_allocs;
    const char* build_base = nullptr;  // rm_upload_scene: address of the DevScene being filled (for SceneAlloc::field_off)
    rm::TreeProgram tree;        // operator-tree scenes: compiled programs (host copy, for the per-frame animation offsets)
    float* d_anim = nullptr;     // device copy of the AnimatedTranslate offsets (inside scene_allocs)
    double anim_time = 0.0;
    bool anim_valid = false;
    // per-launch stats
    rm::DevStats* d_stats = nullptr;
    rm::DevStats* h_stats = nullptr;  // pinned
    rm_stats_t last{};
    // cost-ordered tile queue: per-tile cycle counts of the last frame and the tile order derived from them; reused while the
    // frame geometry (size, band, stripes) stays the same
    rm::DevBuf d_tile_cost, d_tile_order;
    int order_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool order_valid = false;
    // band staging for rm_render (device planes + pinned host mirror)
    rm::DevBuf d_frame, h_frame;
    // early download of finished row bands (rm_render into page-locked planes)
    cudaStream_t copy_stream = nullptr;
    unsigned int* h_band_flags = nullptr;  // page-locked, written by the kernel (kMaxBands words)
    struct EarlyCopy {
        bool on = false;
        int n_bands = 0, band_rows = 0, band_h = 0, width = 0;
        int stripe_rows = 0, stripe_count = 1, stripe_index = 0;  // row stripes: only the owned rows are copied
        struct Plane { char* dst; const char* src; size_t bpp; } planes[9];
        int n_planes = 0;
    } early;
    // user allocations (rm_alloc / rm_host_alloc) and caller memory page-locked by rm_host_register
    std::vector<void*> user_allocs;
    std::vector<std::pair<char*, size_t>> host_allocs;
    std::vector<std::pair<char*, size_t>> host_registered;
    std::vector<std::pair<char*, size_t>> host_shared;  // page-locked by somebody else (the pool): known pinned, not owned
};

namespace rm {
// rm_api.cu internals the pool uses
void free_scene(rm_ctx* c);
}  // namespace rm
