// rm_api.cu — the C ABI declared in include/rm.h: context, scene upload, render, stats, shade, IPC.
//
// One rm_ctx = one CUDA device + one stream.  The scene (primitive arrays, packed fast-path records,
// flattened BVH / octree) is uploaded once and stays resident in HBM; a render call only ships the
// ~140-byte request and brings back the band's planes.  There is no CPU fallback anywhere in this
// file: if the CUDA runtime reports no device, rm_create fails with RM_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "rm_ctx.h"
#include "rm_host.h"
#include "rm_types.h"

using namespace rm;

namespace {
thread_local std::string g_create_error;
}  // namespace

namespace {

constexpr size_t kEarlyCopyMinBytes = 16u << 20;  // frames below this are downloaded in one piece after the kernel
bool early_copy_enabled() {  // RM_EARLY_COPY=0: test / measurement knob, read per call
    const char* e = std::getenv("RM_EARLY_COPY");
    return !(e && e[0] == '0');
}

// Rows of [y0, y1) this request owns, as maximal contiguous spans: everything without row stripes, else the pieces of
// the owned stripes (stripe s covers rows [s * stripe_rows, (s + 1) * stripe_rows), owned when s % count == index).
template <typename F>
void for_owned_spans(int stripe_rows, int stripe_count, int stripe_index, int y0, int y1, F&& fn) {
    if (y1 <= y0) return;
    if (stripe_count <= 1) {
        fn(y0, y1 - y0);
        return;
    }
    for (int sIdx = y0 / stripe_rows; sIdx * stripe_rows < y1; ++sIdx) {
        if (sIdx % stripe_count != stripe_index) continue;
        const int a = std::max(y0, sIdx * stripe_rows), b = std::min(y1, (sIdx + 1) * stripe_rows);
        if (b > a) fn(a, b - a);
    }
}

// D2H of the owned rows of [y0, y1) of every plane.  With row stripes the full stripes of a plane go out as ONE strided
// 2-D copy (pitch = one stripe period), so a band costs one or two copy calls per plane however many stripes it holds.
cudaError_t copy_owned_rows(const rm_ctx::EarlyCopy& ec, int y0, int y1, cudaStream_t s) {
    y1 = std::min(y1, ec.band_h);
    if (y1 <= y0) return cudaSuccess;
    for (int k = 0; k < ec.n_planes; ++k) {
        const size_t rowB = (size_t)ec.width * ec.planes[k].bpp;
        char* dst = ec.planes[k].dst;
        const char* src = ec.planes[k].src;
        if (ec.stripe_count <= 1) {
            cudaError_t e = cudaMemcpyAsync(dst + y0 * rowB, src + y0 * rowB, (size_t)(y1 - y0) * rowB, cudaMemcpyDeviceToHost, s);
            if (e != cudaSuccess) return e;
            continue;
        }
        // full owned stripes whose start rows are one period apart -> one 2-D copy; ragged pieces -> 1-D copies
        int firstFull = -1, nFull = 0;
        cudaError_t err = cudaSuccess;
        const size_t period = (size_t)ec.stripe_rows * ec.stripe_count;
        auto flush = [&]() {
            if (nFull > 0 && err == cudaSuccess)
                err = cudaMemcpy2DAsync(dst + firstFull * rowB, period * rowB, src + firstFull * rowB, period * rowB,
                                        (size_t)ec.stripe_rows * rowB, (size_t)nFull, cudaMemcpyDeviceToHost, s);
            nFull = 0;
            firstFull = -1;
        };
        for_owned_spans(ec.stripe_rows, ec.stripe_count, ec.stripe_index, y0, y1, [&](int a, int rows) {
            if (rows == ec.stripe_rows && (nFull == 0 || (size_t)a == firstFull + nFull * period)) {
                if (nFull == 0) firstFull = a;
                ++nFull;
                return;
            }
            flush();
            if (rows == ec.stripe_rows) {
                firstFull = a;
                nFull = 1;
            } else if (err == cudaSuccess) {
                err = cudaMemcpyAsync(dst + a * rowB, src + a * rowB, (size_t)rows * rowB, cudaMemcpyDeviceToHost, s);
            }
        });
        flush();
        if (err != cudaSuccess) return err;
    }
    return cudaSuccess;
}

int fail(rm_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    else g_create_error = buf;
    return code;
}
#define CU(c, call)                                                                             \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) return fail(c, RM_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

}  // namespace
void rm::free_scene(rm_ctx* c) {
    for (const SceneAlloc& a : c->scene_allocs) cudaFree(a.p);
    c->scene_allocs.clear();
    c->has_scene = false;
    c->exact_only = false;
    c->tree = TreeProgram();
    c->d_anim = nullptr;
    c->anim_valid = false;
    std::memset(&c->scene, 0, sizeof(c->scene));
}
namespace {

template <class T>
int upload(rm_ctx* c, const T* host, size_t n, const T** dev_out) {
    *dev_out = nullptr;
    if (n == 0) return RM_OK;
    void* d = nullptr;
    CU(c, cudaMalloc(&d, n * sizeof(T)));
    c->scene_allocs.push_back({d, n * sizeof(T), (ptrdiff_t)((const char*)dev_out - c->build_base)});
    CU(c, cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    *dev_out = (const T*)d;
    return RM_OK;
}

int ensure(rm_ctx* c, DevBuf& b, size_t bytes, bool pinned_host) {
    if (b.cap >= bytes) return RM_OK;
    if (b.p) {
        if (pinned_host) cudaFreeHost(b.p);
        else cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    }
    size_t cap = bytes + bytes / 8 + 256;
    if (pinned_host) CU(c, cudaMallocHost(&b.p, cap));
    else CU(c, cudaMalloc(&b.p, cap));
    b.cap = cap;
    return RM_OK;
}

bool tile_order_enabled() {  // RM_TILE_ORDER=0: measurement knob (identity queue order), read per call
    const char* e = std::getenv("RM_TILE_ORDER");
    return !(e && e[0] == '0');
}
bool tile_order_forced() {  // RM_TILE_ORDER=2: measurement knob (ordered queue for unstriped requests too)
    const char* e = std::getenv("RM_TILE_ORDER");
    return e && e[0] == '2';
}
constexpr int kTailTriggerWarps = 4;  // end of the frame: a cooperative pass as soon as this many warps are parked (RM_TAIL_TRIGGER overrides)
constexpr int kTileOrderMinTiles = 8192;  // smaller frames finish in a handful of tile rounds: nothing to order

bool is_translation_sphere(uint8_t type, const float* m) {
    if (type != RM_PRIM_SPHERE) return false;
    static const float I[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    for (int i = 0; i < 12; ++i)
        if (m[i] != I[i]) return false;  // -0 == 0
    return m[15] == 1.f;
}
bool is_affine(const float* m) { return m[3] == 0.f && m[7] == 0.f && m[11] == 0.f && m[15] == 1.f; }

int validate_request(rm_ctx* c, const rm_request* rq) {
    if (!rq) return fail(c, RM_ERR_ARG, "request is null");
    if (rq->width <= 0 || rq->height <= 0) return fail(c, RM_ERR_ARG, "bad frame size %dx%d", rq->width, rq->height);
    if (rq->y_start < 0 || rq->y_end > rq->height) return fail(c, RM_ERR_ARG, "band [%d,%d) outside frame height %d", rq->y_start, rq->y_end, rq->height);
    if (rq->algorithm < 0 || rq->algorithm > RM_ALG_ADAPTIVE_STEP_V3) return fail(c, RM_ERR_ARG, "bad algorithm %d", rq->algorithm);
    if (rq->shader < RM_SHADER_NONE || rq->shader > RM_SHADER_ITERATION_HEATMAP) return fail(c, RM_ERR_ARG, "bad shader %d", rq->shader);
    if (rq->shader_analytics < RM_SHADER_NONE || rq->shader_analytics > RM_SHADER_ITERATION_HEATMAP)
        return fail(c, RM_ERR_ARG, "bad analytics shader %d", rq->shader_analytics);
    if (rq->stripe_count > 1 && (rq->stripe_rows <= 0 || rq->stripe_rows % kTileH != 0 || rq->stripe_index < 0 || rq->stripe_index >= rq->stripe_count))
        return fail(c, RM_ERR_ARG, "bad stripe spec rows=%d count=%d index=%d (rows must be a positive multiple of %d)", rq->stripe_rows, rq->stripe_count, rq->stripe_index, kTileH);
    if (rq->stripe_count < 0) return fail(c, RM_ERR_ARG, "bad stripe_count %d", rq->stripe_count);
    return RM_OK;
}

// Enqueue the render of one band into device planes; read the stats back; fill c->last.
int render_device_locked(rm_ctx* c, const rm_request* rq, const rm_result* out, cudaStream_t stream) {
    auto w0 = std::chrono::steady_clock::now();
    if (!c->has_scene) return fail(c, RM_ERR_STATE, "rm_render before rm_upload_scene");
    int rc = validate_request(c, rq);
    if (rc) return rc;
    if (!out || !out->depth || !out->normal || !out->sdf_eval || !out->iters)
        return fail(c, RM_ERR_ARG, "result planes depth/normal/sdf_eval/iters are required");
    const int bandH = rq->y_end > rq->y_start ? rq->y_end - rq->y_start : 0;  // Math.max(0, yEnd - yStart)
    CU(c, cudaSetDevice(c->device));

    RenderParams P{};
    P.scene = c->scene;
    P.width = rq->width;
    P.height = rq->height;
    P.y_start = rq->y_start;
    P.y_end = rq->y_start + bandH;
    std::memcpy(P.rot3, rq->rot3, sizeof(P.rot3));
    std::memcpy(P.origin, rq->origin, sizeof(P.origin));
    P.algorithm = rq->algorithm;
    P.step_size = rq->step_size;
    P.overshoot = rq->overshoot_factor;
    P.shader = rq->shader;
    P.shader2 = rq->shader_analytics;
    P.length_sqrt = (c->flags & RM_F_LENGTH_SQRT) ? 1 : 0;
    {
        const char* a = std::getenv("RM_ANATOMY");  // measurement knob: drain_ms / tail_ms of rm_stats (off: no timestamps taken)
        P.anatomy = (a && a[0] == '1') ? 1 : 0;
        const char* e = std::getenv("RM_TAIL_TRIGGER");  // measurement knob, read per call
        P.tail_trigger = e ? std::max(1, std::atoi(e)) : kTailTriggerWarps;
    }
    const bool tree = c->exact_only;
    P.scene.time = rq->time;
    // Operator trees always run the exact (fp64, unfused) kernels: B200 has full-rate-class fp64 and these scenes
    // are a handful of primitives, so there is no reduced-precision variant to disagree with the reference.  A
    // context created without RM_F_VALIDATE_FP64 only swaps V8's compensated hypot for a plain sqrt.
    if (tree && !(c->flags & RM_F_VALIDATE_FP64)) {
        P.length_sqrt = 1;
        const char* e = std::getenv("RM_FAST_OBJECTS");  // RM_FAST_OBJECTS=0: measurement knob (fp64 object evaluators everywhere)
        P.fast_objects = (e && e[0] == '0') ? 0 : 1;
    }
    if (tree && !c->tree.anims.empty() && (!c->anim_valid || c->anim_time != rq->time)) {
        // Scene.updateTime (raymarcher.ts:59): AnimatedTranslate's offset vector is a per-job constant
        std::vector<float> off;
        eval_anim_offsets(c->tree, rq->time, off);
        CU(c, cudaMemcpyAsync(c->d_anim, off.data(), off.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
        CU(c, cudaStreamSynchronize(stream));  // `off` is pageable and goes out of scope
        c->anim_time = rq->time;
        c->anim_valid = true;
    }
    P.tiles_x = (rq->width + kTileW - 1) / kTileW;
    uint64_t ownedRows = (uint64_t)bandH;
    if (rq->stripe_count > 1) {
        P.stripe_rows = rq->stripe_rows;
        P.stripe_count = rq->stripe_count;
        P.stripe_index = rq->stripe_index;
        const int nStripes = (bandH + P.stripe_rows - 1) / P.stripe_rows;
        const int nOwned = nStripes > P.stripe_index ? (nStripes - P.stripe_index + P.stripe_count - 1) / P.stripe_count : 0;
        P.tiles_per_stripe = P.tiles_x * (P.stripe_rows / kTileH);
        P.n_tiles = nOwned * P.tiles_per_stripe;
        ownedRows = 0;
        for (int sIdx = P.stripe_index; sIdx < nStripes; sIdx += P.stripe_count) {
            int y0 = sIdx * P.stripe_rows, y1 = y0 + P.stripe_rows < bandH ? y0 + P.stripe_rows : bandH;
            ownedRows += (uint64_t)(y1 - y0);
        }
    } else {
        P.stripe_rows = kTileH;
        P.stripe_count = 1;
        P.stripe_index = 0;
        P.tiles_per_stripe = P.tiles_x;
        P.n_tiles = P.tiles_x * ((bandH + kTileH - 1) / kTileH);
    }
    P.depth = out->depth;
    P.normal = out->normal;
    P.sdf = out->sdf_eval;
    P.iters = out->iters;
    P.rgba = (rq->shader != RM_SHADER_NONE) ? out->rgba : nullptr;
    P.rgba2 = (rq->shader_analytics != RM_SHADER_NONE) ? out->rgba_analytics : nullptr;
    P.depth_f32 = out->depth_f32;
    P.sdf_u32 = out->sdf_eval_u32;
    P.depth_f64 = out->depth_f64;
    {
        auto al16 = [](const void* q) { return ((uintptr_t)q & 15u) == 0; };
        P.vec_store = (rq->width % kTileW == 0 && al16(P.depth) && al16(P.normal) && al16(P.sdf) && al16(P.iters) && al16(P.rgba) && al16(P.rgba2)) ? 1 : 0;
    }
    P.stats = c->d_stats;
    // cost-ordered queue: hand the tiles out most-expensive-first, by what each tile cost in the previous frame of this geometry
    // (only where tile costs are heavy-tailed: the cooperative-queue kernels of big BVH scenes.  On cheap, uniform frames the
    // reordering loses more in locality than the shorter tail gains: cfg1 +4 %, cfg5 +2 %, profiles/r02g_ab.jsonl)
    // Measured on cfg4 (profiles/r02k_stripe_time.log): the 1/2, 1/4, 1/8 stripe shares of a multi-GPU frame gain 3-5 %, the whole
    // frame on one GPU loses 2 % (29.0 -> 29.7 ms): the order is used for striped requests only.
    const bool ordered = P.n_tiles >= kTileOrderMinTiles && tile_order_enabled() && !(c->flags & RM_F_VALIDATE_FP64) && !tree &&
                         c->scene.accel_kind == RM_ACCEL_BVH && c->scene.n_prims >= 256 && (rq->stripe_count > 1 || tile_order_forced());
    if (ordered) {
        const int key[8] = {rq->width, rq->height, rq->y_start, bandH, P.stripe_rows, P.stripe_count, P.stripe_index, P.n_tiles};
        const size_t bytes = (size_t)P.n_tiles * sizeof(unsigned int);
        if (c->d_tile_cost.cap < bytes || c->d_tile_order.cap < bytes) c->order_valid = false;
        if ((rc = ensure(c, c->d_tile_cost, bytes, false)) || (rc = ensure(c, c->d_tile_order, bytes, false))) return rc;
        if (std::memcmp(key, c->order_key, sizeof(key)) != 0) c->order_valid = false;
        std::memcpy(c->order_key, key, sizeof(key));
        P.tile_cost = (unsigned int*)c->d_tile_cost.p;
        P.tile_order = c->order_valid ? (const unsigned int*)c->d_tile_order.p : nullptr;
    }
    const bool early = c->early.on && P.n_tiles > 0;
    if (early) {
        P.band_flags = c->h_band_flags;
        P.band_rows = c->early.band_rows;
        for (int b = 0; b < kMaxBands; ++b) {
            c->h_band_flags[b] = 0u;
            uint64_t rows = 0;
            for_owned_spans(P.stripe_rows, P.stripe_count, P.stripe_index, b * P.band_rows, std::min((b + 1) * P.band_rows, bandH),
                            [&](int, int n) { rows += (uint64_t)n; });
            P.band_px[b] = b < c->early.n_bands ? (uint32_t)(rows * (uint64_t)rq->width) : 0u;
        }
    }

    DevStats init{};
    init.min_sdf = 0xffffffffu;
    init.min_iters = 0xffffffffu;
    init.t_enter_min = init.t_drain_min = init.t_exit_min = ~0ull;
    *c->h_stats = init;
    CU(c, cudaMemcpyAsync(c->d_stats, c->h_stats, sizeof(DevStats), cudaMemcpyHostToDevice, stream));
    int launches = 0;
    CU(c, cudaEventRecord(c->ev0, stream));
    if (P.n_tiles > 0) {
        int e = ((c->flags & RM_F_VALIDATE_FP64) || tree) ? launch_render_val(P, c->n_sms, stream) : launch_render_fast(P, c->n_sms, stream);
        if (e != 0) return fail(c, RM_ERR_CUDA, "render launch: %s", cudaGetErrorString((cudaError_t)e));
        launches = 1;
    }
    CU(c, cudaEventRecord(c->ev1, stream));
    CU(c, cudaMemcpyAsync(c->h_stats, c->d_stats, sizeof(DevStats), cudaMemcpyDeviceToHost, stream));
    if (ordered && launches) {  // next frame's order from this frame's per-tile costs (one small CTA, after the timed kernel)
        // spatial runs of >= 4 tiles per resident warp (16 warps per SM on the kernels that use the order), at most 16
        // (frames whose bands are not downloaded during the render need no spatial order: 0 = every tile by cost)
        const int runs = early ? std::max(1, std::min(16, P.n_tiles / (4 * 16 * c->n_sms))) : 0;
        int e = launch_order_tiles((const unsigned int*)c->d_tile_cost.p, (unsigned int*)c->d_tile_order.p, P.n_tiles, runs, stream);
        if (e != 0) return fail(c, RM_ERR_CUDA, "tile-order launch: %s", cudaGetErrorString((cudaError_t)e));
        c->order_valid = true;
        launches = 2;
    }
    if (early) {
        // download each row band as soon as the kernel reports it finished; what is left goes out after the kernel
        const rm_ctx::EarlyCopy& ec = c->early;
        bool sent[kMaxBands] = {};
        auto send = [&](int b) -> cudaError_t {
            sent[b] = true;
            return copy_owned_rows(ec, b * ec.band_rows, (b + 1) * ec.band_rows, c->copy_stream);
        };
        for (;;) {
            const cudaError_t q = cudaStreamQuery(stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) CU(c, q);
            for (int b = 0; b < ec.n_bands; ++b)
                if (!sent[b] && (q == cudaSuccess || *(volatile unsigned int*)&c->h_band_flags[b])) CU(c, send(b));
            if (q == cudaSuccess) break;
        }
        CU(c, cudaStreamSynchronize(c->copy_stream));
    }
    CU(c, cudaStreamSynchronize(stream));
    float ms = 0.f;
    CU(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));

    const DevStats& s = *c->h_stats;
    if (s.t_total) fprintf(stderr, "[rm phase timing] warp-cycles total %.3e search %.1f%% barrier %.1f%% stuck %.1f%%\n", (double)s.t_total, 100.0 * s.t_search / s.t_total, 100.0 * s.t_barrier / s.t_total, 100.0 * s.t_stuck / s.t_total);
    if (s.n_pass) fprintf(stderr, "[rm phase timing] passes %llu, requests/pass %.1f, tc_pass cycles/pass (thread 0): sweep %.0f publish %.0f drain %.0f setup %.0f; work items per pass: %.1f\n", s.n_pass, (double)s.n_req / s.n_pass, (double)s.t_tc[0] / s.n_pass, (double)s.t_tc[1] / s.n_pass, (double)s.t_tc[2] / s.n_pass, (double)s.t_tc[3] / s.n_pass, (double)s.n_rearm / s.n_pass);
    rm_stats_t& L = c->last;
    std::memset(&L, 0, sizeof(L));
    L.n_pixels = (uint64_t)rq->width * ownedRows;
    L.sum_sdf = s.sum_sdf;
    L.sum_iters = s.sum_iters;
    L.max_sdf = s.max_sdf;
    L.min_sdf = s.min_sdf;
    L.max_iters = s.max_iters;
    L.min_iters = s.min_iters;
    L.sum_sdf_full = s.sum_sdf_full;
    L.sum_iters_full = s.sum_iters_full;
    L.evals_by_type[0] = s.evals_sphere;
    L.evals_by_type[1] = s.evals_box;
    L.evals_by_type[2] = s.evals_torus;
    L.n_hit = s.n_hit;
    {
        const bool ts = !(c->flags & RM_F_VALIDATE_FP64) && c->scene.prim_kind == PK_TSPHERE;
        // translation-only spheres: 11 FLOP (3 add, 3 mul, 2 add, sqrt, sub, min); the screened search of large scenes
        // executes 7 per primitive (2 p.t + |t|^2 as 3 FMA, then min; no sqrt / radius subtraction in the hot loop)
        const double tsFlops = c->scene.n_chunks >= 16 ? 7.0 : 11.0;
        L.operator_flops = (double)s.op_flops;
        L.algorithmic_flops = ts ? tsFlops * (double)s.evals_sphere
                                 : 26.0 * (double)s.evals_sphere + 38.0 * (double)s.evals_box + 29.0 * (double)s.evals_torus + L.operator_flops;
    }
    L.tc_passes = s.tc_passes;
    L.tc_requests = s.tc_requests;
    L.tc_items = s.tc_items;
    L.fp32_pipe_flops = L.algorithmic_flops;  // every counted evaluation was performed (validation build: in fp64)
    L.tensor_flops = 0.0;
    if (s.tc_passes) {
        // leaf evaluations (11 FLOP) + 128 spheres per work item on the FP32 pipe; two tf32 sweeps of 128 x 128 x 16 MACs per
        // cluster block on the tensor cores
        const double leafEvals = (double)s.evals_sphere - (double)s.tc_requests * (double)c->scene.n_prims;
        L.fp32_pipe_flops = 11.0 * leafEvals + (double)s.tc_items * 128.0 * 11.0;
        L.tensor_flops = (double)s.tc_passes * (double)c->scene.n_tc_blocks * 2.0 * (2.0 * 128 * 128 * 16);
    }
    L.executed_flops = L.fp32_pipe_flops + L.tensor_flops;
    L.n_devices = 1;
    if (s.t_exit_max && s.t_enter_min != ~0ull && s.t_drain_min != ~0ull) {
        // frame anatomy: first CTA in -> first warp finds the tile queue empty (steady state) -> last warp out (tail)
        L.drain_ms = (double)(s.t_drain_min - s.t_enter_min) * 1e-6;
        L.tail_ms = (double)(s.t_exit_max - s.t_drain_min) * 1e-6;
    }
    L.kernel_ms = ms;
    L.n_launches = launches;
    L.device = c->device;
    L.wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    return RM_OK;
}

}  // namespace

extern "C" {

int rm_abi_version(void) { return RM_ABI_VERSION; }

int rm_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e);
        return RM_ERR_CUDA;
    }
    return n;
}

int rm_create(rm_ctx** out, int device, unsigned flags) {
    if (!out) return fail(nullptr, RM_ERR_ARG, "out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(nullptr, RM_ERR_CUDA, "no CUDA device available (%s); librm_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= n) return fail(nullptr, RM_ERR_ARG, "device %d out of range [0,%d)", device, n);
    rm_ctx* c = new (std::nothrow) rm_ctx();
    if (!c) return fail(nullptr, RM_ERR_NOMEM, "out of memory");
    c->device = device;
    c->flags = flags;
#define CUC(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e2 = (call);                                                                         \
        if (e2 != cudaSuccess) {                                                                         \
            int rc_ = fail(nullptr, RM_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e2));               \
            rm_destroy(c);                                                                               \
            return rc_;                                                                                  \
        }                                                                                                \
    } while (0)
    CUC(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUC(cudaGetDeviceProperties(&prop, device));
    c->n_sms = prop.multiProcessorCount;
    CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUC(cudaEventCreate(&c->ev0));
    CUC(cudaEventCreate(&c->ev1));
    CUC(cudaMalloc((void**)&c->d_stats, sizeof(DevStats)));
    CUC(cudaMallocHost((void**)&c->h_stats, sizeof(DevStats)));
    CUC(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CUC(cudaHostAlloc((void**)&c->h_band_flags, kMaxBands * sizeof(unsigned int), cudaHostAllocMapped));
#undef CUC
    *out = c;
    return RM_OK;
}

void rm_destroy(rm_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    free_scene(c);
    for (void* p : c->user_allocs) cudaFree(p);
    for (auto& h : c->host_allocs) cudaFreeHost(h.first);
    for (auto& h : c->host_registered) cudaHostUnregister(h.first);
    if (c->d_tile_cost.p) cudaFree(c->d_tile_cost.p);
    if (c->d_tile_order.p) cudaFree(c->d_tile_order.p);
    if (c->d_frame.p) cudaFree(c->d_frame.p);
    if (c->h_frame.p) cudaFreeHost(c->h_frame.p);
    if (c->d_stats) cudaFree(c->d_stats);
    if (c->h_stats) cudaFreeHost(c->h_stats);
    if (c->h_band_flags) cudaFreeHost(c->h_band_flags);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* rm_last_error(rm_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int rm_build_bvh(int32_t n, const uint8_t* type, const float* w2l, const double* params, unsigned flags, rm_bvh_node* nodes,
                 int32_t* n_nodes, int32_t* leaf, int32_t* n_leaf) {
    if (n < 0 || (n > 0 && (!type || !w2l || !params)) || !n_nodes || !n_leaf) return RM_ERR_ARG;
    std::vector<PrimGeom> geom;
    compute_prim_geometry(n, type, w2l, params, flags, geom);
    std::vector<rm_bvh_node> nn;
    std::vector<int32_t> lp;
    build_bvh(geom, nn, lp);
    if (nodes) {
        if (*n_nodes < (int32_t)nn.size() || *n_leaf < (int32_t)lp.size() || (!leaf && !lp.empty())) return RM_ERR_ARG;
        std::memcpy(nodes, nn.data(), nn.size() * sizeof(rm_bvh_node));
        if (!lp.empty()) std::memcpy(leaf, lp.data(), lp.size() * sizeof(int32_t));
    }
    *n_nodes = (int32_t)nn.size();
    *n_leaf = (int32_t)lp.size();
    return RM_OK;
}

int rm_build_octree(int32_t n, const uint8_t* type, const float* w2l, const double* params, unsigned flags,
                    rm_octree_node* nodes, int32_t* n_nodes, int32_t* leaf, int32_t* n_leaf) {
    if (n < 0 || (n > 0 && (!type || !w2l || !params)) || !n_nodes || !n_leaf) return RM_ERR_ARG;
    std::vector<PrimGeom> geom;
    compute_prim_geometry(n, type, w2l, params, flags, geom);
    std::vector<rm_octree_node> nn;
    std::vector<int32_t> lp;
    build_octree(geom, nn, lp);
    if (nodes) {
        if (*n_nodes < (int32_t)nn.size() || *n_leaf < (int32_t)lp.size() || (!leaf && !lp.empty())) return RM_ERR_ARG;
        std::memcpy(nodes, nn.data(), nn.size() * sizeof(rm_octree_node));
        if (!lp.empty()) std::memcpy(leaf, lp.data(), lp.size() * sizeof(int32_t));
    }
    *n_nodes = (int32_t)nn.size();
    *n_leaf = (int32_t)lp.size();
    return RM_OK;
}

static int check_scene_arrays(const rm_scene* s) {
    if (!s || s->n_prims < 0 || (s->n_prims > 0 && (!s->type || !s->world_to_local || !s->params))) return RM_ERR_ARG;
    for (int32_t i = 0; i < s->n_prims; ++i)
        if (s->type[i] > RM_PRIM_MANDELBULB) return RM_ERR_UNSUPPORTED_PRIMITIVE;
    std::string err;
    return validate_tree(*s, err);
}

int rm_build_bvh_scene(const rm_scene* s, unsigned flags, rm_bvh_node* nodes, int32_t* n_nodes, int32_t* leaf, int32_t* n_leaf) {
    if (!n_nodes || !n_leaf) return RM_ERR_ARG;
    int rc = check_scene_arrays(s);
    if (rc) return rc;
    std::vector<PrimGeom> geom;
    compute_object_geometry(*s, flags, geom);
    std::vector<rm_bvh_node> nn;
    std::vector<int32_t> lp;
    build_bvh(geom, nn, lp);
    if (nodes) {
        if (*n_nodes < (int32_t)nn.size() || *n_leaf < (int32_t)lp.size() || (!leaf && !lp.empty())) return RM_ERR_ARG;
        std::memcpy(nodes, nn.data(), nn.size() * sizeof(rm_bvh_node));
        if (!lp.empty()) std::memcpy(leaf, lp.data(), lp.size() * sizeof(int32_t));
    }
    *n_nodes = (int32_t)nn.size();
    *n_leaf = (int32_t)lp.size();
    return RM_OK;
}

int rm_build_octree_scene(const rm_scene* s, unsigned flags, rm_octree_node* nodes, int32_t* n_nodes, int32_t* leaf, int32_t* n_leaf) {
    if (!n_nodes || !n_leaf) return RM_ERR_ARG;
    int rc = check_scene_arrays(s);
    if (rc) return rc;
    std::vector<PrimGeom> geom;
    compute_object_geometry(*s, flags, geom);
    std::vector<rm_octree_node> nn;
    std::vector<int32_t> lp;
    build_octree(geom, nn, lp);
    if (nodes) {
        if (*n_nodes < (int32_t)nn.size() || *n_leaf < (int32_t)lp.size() || (!leaf && !lp.empty())) return RM_ERR_ARG;
        std::memcpy(nodes, nn.data(), nn.size() * sizeof(rm_octree_node));
        if (!lp.empty()) std::memcpy(leaf, lp.data(), lp.size() * sizeof(int32_t));
    }
    *n_nodes = (int32_t)nn.size();
    *n_leaf = (int32_t)lp.size();
    return RM_OK;
}

int rm_upload_scene(rm_ctx* c, const rm_scene* s) {
    if (!c) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!s) return fail(c, RM_ERR_ARG, "scene is null");
    const int32_t n = s->n_prims;
    if (n < 0 || (n > 0 && (!s->type || !s->world_to_local || !s->params))) return fail(c, RM_ERR_ARG, "bad primitive arrays");
    if (s->accel_kind < RM_ACCEL_NONE || s->accel_kind > RM_ACCEL_BVH) return fail(c, RM_ERR_ARG, "bad accel_kind %d", s->accel_kind);
    // Primitives: sphere / box / torus; scene objects: those, or operator trees over them.
    // Anything else (mandelbulb, unknown kinds) is rejected loudly.
    const bool tree = s->n_objects > 0;
    bool hasBulb = false;
    for (int32_t i = 0; i < n; ++i) hasBulb = hasBulb || s->type[i] == RM_PRIM_MANDELBULB;
    const bool exactOnly = tree || hasBulb;  // these scenes run the exact (fp64) kernels in every context
    {
        std::string terr;
        int trc = validate_tree(*s, terr);
        if (trc) return fail(c, trc, "%s", terr.c_str());
    }
    const int32_t nObj = tree ? s->n_objects : n;  // Scene.objectSDFs.length
    bool allTS = n > 0 && !exactOnly;
    uint32_t hist[3] = {0, 0, 0};
    for (int32_t i = 0; i < n; ++i) {
        if (s->type[i] > RM_PRIM_MANDELBULB)
            return fail(c, RM_ERR_UNSUPPORTED_PRIMITIVE, "primitive %d has type %d: only sphere/box/torus/mandelbulb are supported (no CPU fallback)", i, (int)s->type[i]);
        if (s->type[i] == RM_PRIM_MANDELBULB) {
            const double* q = s->params + 4 * (size_t)i;
            if (!(q[1] >= 0.0 && q[1] <= 100000.0)) return fail(c, RM_ERR_ARG, "primitive %d: mandelbulb iteration count %g out of range", i, q[1]);
        } else {
            hist[s->type[i]]++;
        }
        const float* m = s->world_to_local + 16 * (size_t)i;
        for (int k = 0; k < 16; ++k)
            if (!std::isfinite(m[k])) return fail(c, RM_ERR_ARG, "primitive %d: non-finite transform", i);
        if (!(c->flags & RM_F_VALIDATE_FP64) && !exactOnly && !is_affine(m))
            return fail(c, RM_ERR_UNSUPPORTED_PRIMITIVE, "primitive %d: projective world->local is only supported by the validation build", i);
        allTS = allTS && is_translation_sphere(s->type[i], m);
    }
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    free_scene(c);
    c->order_valid = false;  // tile costs belong to the old scene

    DevScene ds{};
    c->build_base = (const char*)&ds;  // every upload() below fills a pointer field of `ds`
    ds.n_prims = nObj;
    ds.accel_kind = s->accel_kind;
    ds.prim_kind = allTS ? PK_TSPHERE : PK_GENERAL;
    std::memcpy(ds.type_hist, hist, sizeof(hist));
    int rc;
    if ((rc = upload(c, s->type, (size_t)n, &ds.type))) return rc;
    if ((rc = upload(c, s->world_to_local, (size_t)n * 16, &ds.w2l))) return rc;
    if ((rc = upload(c, s->params, (size_t)n * 4, &ds.params))) return rc;

    if (tree) {
        TreeProgram& tp = c->tree;
        if (!compile_tree(*s, !(c->flags & RM_F_VALIDATE_FP64), tp))
            return fail(c, RM_ERR_ARG, "an operator tree has more than 255 leaves of one primitive type");
        ds.n_instrs = (int32_t)tp.instrs.size();
        if ((rc = upload(c, tp.instrs.data(), tp.instrs.size(), &ds.instrs))) return rc;
        if ((rc = upload(c, tp.obj_first.data(), tp.obj_first.size(), &ds.obj_first))) return rc;
        if ((rc = upload(c, tp.mats.data(), tp.mats.size(), &ds.mats))) return rc;
        if ((rc = upload(c, tp.obj_hist.data(), tp.obj_hist.size(), &ds.obj_hist))) return rc;
        if ((rc = upload(c, tp.obj_flops.data(), tp.obj_flops.size(), &ds.obj_flops))) return rc;
        for (int k = 0; k < 3; ++k) ds.type_hist[k] = 0;  // leaf evaluations of one pass over every object
        for (uint32_t h : tp.obj_hist) {
            ds.type_hist[0] += h & 255u;
            ds.type_hist[1] += (h >> 8) & 255u;
            ds.type_hist[2] += (h >> 16) & 255u;
        }
        for (uint32_t f : tp.obj_flops) ds.all_op_flops += f;
        if (!tp.anims.empty()) {
            std::vector<float> off;
            eval_anim_offsets(tp, 0.0, off);
            if ((rc = upload(c, off.data(), off.size(), &ds.anim))) return rc;
            CU(c, cudaStreamSynchronize(c->stream));
            c->d_anim = const_cast<float*>(ds.anim);
            c->anim_time = 0.0;
            c->anim_valid = true;
        }
    }

    // fast-path packed records (plain primitive lists only)
    std::vector<float4> rec;
    if (exactOnly) {
    } else if (allTS) {
        const int32_t nChunks = (n + 31) / 32;
        rec.resize((size_t)nChunks * 40);  // 40 float4 = 160 floats = tx[32] ty[32] tz[32] r[32] tt[32]
        float* f = reinterpret_cast<float*>(rec.data());
        float rmin = 3.0e38f, rmax = -3.0e38f, ttmax = 0.f;
        std::vector<float> chunkRmax(((size_t)nChunks + 3) / 4 * 4, 0.f);  // padded: the kernel reads one float4 per stage
        for (int32_t i = 0; i < nChunks * 32; ++i) {
            float* c = f + (size_t)(i >> 5) * 160 + (i & 31);
            if (i < n) {
                const float* m = s->world_to_local + 16 * (size_t)i;
                const float r = (float)s->params[4 * (size_t)i];
                c[0] = m[12];
                c[32] = m[13];
                c[64] = m[14];
                c[96] = r;
                c[128] = (float)((double)m[12] * m[12] + (double)m[13] * m[13] + (double)m[14] * m[14]);
                ttmax = std::fmax(ttmax, c[128]);
                rmin = std::fmin(rmin, r);
                rmax = std::fmax(rmax, r);
                chunkRmax[(size_t)(i >> 5)] = std::fmax(chunkRmax[(size_t)(i >> 5)], r);
            } else {  // padding: a sphere of radius 0 so far away that it can never be the minimum
                c[0] = c[32] = c[64] = 1.0e15f;
                c[96] = 0.f;
                c[128] = 3.0e30f;
            }
        }
        ds.n_chunks = nChunks;
        ds.r_min = rmin;
        ds.r_max = rmax;
        ds.tt_max = ttmax;
        std::vector<float4> one((size_t)n);
        for (int32_t i = 0; i < n; ++i) {
            const float* m = s->world_to_local + 16 * (size_t)i;
            one[(size_t)i] = make_float4(m[12], m[13], m[14], (float)s->params[4 * (size_t)i]);
        }
        // Cluster screen of the cooperative pass (rm_device.cuh tc_pass).  The spheres are put in a balanced kd order and cut
        // into clusters of 128 (= four chunks of a second, sorted chunk-SoA copy).  Each cluster j gets a centre C_j and two bounds
        //   R_j = max_i(|c_i - C_j| + r_i)   ->  every member's SDF at p is >= |p - C_j| - R_j
        //   u_j = min_i(|c_i - C_j| - r_i)   ->  some member's SDF at p is  <= |p - C_j| + u_j
        // The kernel computes |p - C_j|^2 for 128 queries x all clusters on the tensor cores (split-TF32 B tiles, pre-tiled as
        // shared-memory images) and evaluates only the clusters whose lower
        // bound does not exceed the best upper bound (work items spread over the CTA's warps).  The minimum over ALL spheres is unchanged.
        // (work items pack the cluster index in 16 bits: scenes beyond 65 535 clusters = 8.3 M spheres keep the FFMA search)
        if (n >= 256 && (n + 127) / 128 <= 65535 && s->accel_kind == RM_ACCEL_BVH && !(c->flags & RM_F_VALIDATE_FP64) &&
            !std::getenv("RM_DISABLE_TC")) {
            auto rna = [](float x) {  // cvt.rna.tf32.f32
                uint32_t u;
                std::memcpy(&u, &x, 4);
                u = (u + 0x1000u) & ~0x1FFFu;
                float r;
                std::memcpy(&r, &u, 4);
                return r;
            };
            // Balanced kd ordering of the centres: split the longest axis at the position that keeps every leaf a whole
            // cluster; consecutive runs of kClSize spheres are then compact boxes.
            std::vector<std::pair<uint32_t, int32_t>> order((size_t)n);
            for (int32_t i = 0; i < n; ++i) order[(size_t)i] = {0u, i};
            {
                constexpr int32_t kLeaf = 128;
                std::vector<std::pair<int32_t, int32_t>> stack{{0, n}};
                while (!stack.empty()) {
                    const auto [lo_, hi_] = stack.back();
                    stack.pop_back();
                    if (hi_ - lo_ <= kLeaf) continue;
                    float bl[3] = {3e38f, 3e38f, 3e38f}, bh[3] = {-3e38f, -3e38f, -3e38f};
                    for (int32_t i = lo_; i < hi_; ++i)
                        for (int k = 0; k < 3; ++k) {
                            const float v = s->world_to_local[16 * (size_t)order[(size_t)i].second + 12 + k];
                            bl[k] = std::fmin(bl[k], v);
                            bh[k] = std::fmax(bh[k], v);
                        }
                    int ax = 0;
                    if (bh[1] - bl[1] > bh[ax] - bl[ax]) ax = 1;
                    if (bh[2] - bl[2] > bh[ax] - bl[ax]) ax = 2;
                    const int32_t leaves = (hi_ - lo_ + kLeaf - 1) / kLeaf, mid = lo_ + (leaves / 2) * kLeaf;
                    std::nth_element(order.begin() + lo_, order.begin() + mid, order.begin() + hi_, [&](const auto& a, const auto& b2) {
                        const float va = s->world_to_local[16 * (size_t)a.second + 12 + ax], vb = s->world_to_local[16 * (size_t)b2.second + 12 + ax];
                        return va < vb || (va == vb && a.second < b2.second);
                    });
                    stack.push_back({lo_, mid});
                    stack.push_back({mid, hi_});
                }
            }
            constexpr int32_t kClSize = 128;  // spheres per cluster = 4 chunks of the sorted chunk-SoA copy
            const int32_t nCl = (n + kClSize - 1) / kClSize, nBlocks = (nCl + 127) / 128, nChunksS = nCl * (kClSize / 32);
            std::vector<int32_t> perm((size_t)nChunksS * 32, 0);
            std::vector<float4> clRec((size_t)nChunksS * 40);  // sorted chunk-SoA copy (same layout as `rec`)
            float* cf = reinterpret_cast<float*>(clRec.data());
            for (int32_t i = 0; i < nChunksS * 32; ++i) {
                float* cc = cf + (size_t)(i >> 5) * 160 + (i & 31);
                if (i < n) {
                    const int32_t o = order[(size_t)i].second;
                    perm[(size_t)i] = o;
                    const float* m = s->world_to_local + 16 * (size_t)o;
                    cc[0] = m[12];
                    cc[32] = m[13];
                    cc[64] = m[14];
                    cc[96] = (float)s->params[4 * (size_t)o];
                    cc[128] = (float)((double)m[12] * m[12] + (double)m[13] * m[13] + (double)m[14] * m[14]);
                } else {  // padding: a sphere of radius 0 so far away that it can never be the minimum
                    cc[0] = cc[32] = cc[64] = 1.0e15f;
                    cc[96] = 0.f;
                    cc[128] = 3.0e30f;
                }
            }
            constexpr size_t kBlockFloats = 1024;  // 4 KB tile
            std::vector<float> tiles((size_t)nBlocks * kBlockFloats, 0.f);
            std::vector<float2> bounds((size_t)nBlocks * 128);  // (R_j, u_j)
            std::vector<float> blockRmax((size_t)nBlocks, 0.f);   // largest R_j of each 128-cluster block
            for (int32_t j = 0; j < nBlocks * 128; ++j) {
                float* blk = tiles.data() + (size_t)(j >> 7) * kBlockFloats;
                const int jr = j & 127;
                double C[3] = {-1.0e15, -1.0e15, -1.0e15};  // padding cluster: infinitely far away
                float R = 0.f, U = 0.f;
                if (j < nCl) {
                    const int32_t i0 = j * kClSize, i1 = std::min(n, i0 + kClSize);
                    double bl[3] = {1e300, 1e300, 1e300}, bh[3] = {-1e300, -1e300, -1e300};
                    for (int32_t i = i0; i < i1; ++i) {
                        const int32_t o = order[(size_t)i].second;
                        for (int k = 0; k < 3; ++k) {
                            const double ck = -(double)s->world_to_local[16 * (size_t)o + 12 + k];  // centre = -translation
                            bl[k] = std::min(bl[k], ck);
                            bh[k] = std::max(bh[k], ck);
                        }
                    }
                    for (int k = 0; k < 3; ++k) C[k] = (double)(float)(0.5 * (bl[k] + bh[k]));
                    double Rd = -1e300, Ud = 1e300;
                    for (int32_t i = i0; i < i1; ++i) {
                        const int32_t o = order[(size_t)i].second;
                        double d2 = 0;
                        for (int k = 0; k < 3; ++k) {
                            const double dk = -(double)s->world_to_local[16 * (size_t)o + 12 + k] - C[k];
                            d2 += dk * dk;
                        }
                        const double d = std::sqrt(d2), r = (double)(float)s->params[4 * (size_t)o];
                        Rd = std::max(Rd, d + r);
                        Ud = std::min(Ud, d - r);
                    }
                    // conservative roundings: the bounds are used against fp32 SDF values
                    R = std::nextafter((float)(Rd * (1.0 + 1e-6) + 1e-7), 3e38f);
                    U = std::nextafter((float)(Ud + std::fabs(Ud) * 1e-6 + 1e-7), 3e38f);
                }
                const float t[3] = {(float)-C[0], (float)-C[1], (float)-C[2]};  // "translation" of the cluster centre
                const float tt = std::fmin((float)(C[0] * C[0] + C[1] * C[1] + C[2] * C[2]), 3.0e30f);
                auto at = [&](int k) -> float& { return blk[(k >> 2) * 512 + (jr >> 3) * 32 + (jr & 7) * 4 + (k & 3)]; };
                for (int k = 0; k < 3; ++k) {
                    const float h = rna(t[k]);
                    at(k) = h;
                    at(4 + k) = rna(t[k] - h);
                }
                const float h = rna(tt);
                at(3) = h;
                at(7) = rna(tt - h);
                bounds[(size_t)j] = make_float2(R, U);
                blockRmax[(size_t)(j >> 7)] = std::fmax(blockRmax[(size_t)(j >> 7)], R);
            }
            if ((rc = upload(c, tiles.data(), tiles.size(), &ds.tc_tiles))) return rc;
            if ((rc = upload(c, bounds.data(), bounds.size(), &ds.cl_bound))) return rc;
            if ((rc = upload(c, blockRmax.data(), blockRmax.size(), &ds.cl_block_rmax))) return rc;
            if ((rc = upload(c, clRec.data(), clRec.size(), &ds.cl_rec))) return rc;
            if ((rc = upload(c, perm.data(), perm.size(), &ds.cl_perm))) return rc;
            CU(c, cudaStreamSynchronize(c->stream));  // host vectors go out of scope
            ds.n_tc_blocks = nBlocks;
            ds.n_clusters = nCl;
        }
        if ((rc = upload(c, one.data(), one.size(), &ds.rec1))) return rc;
        if ((rc = upload(c, chunkRmax.data(), chunkRmax.size(), &ds.chunk_rmax))) return rc;
        CU(c, cudaStreamSynchronize(c->stream));  // `one` goes out of scope
    } else {
        rec.resize((size_t)n * 4);
        for (int32_t i = 0; i < n; ++i) {
            const float* m = s->world_to_local + 16 * (size_t)i;
            const double* q = s->params + 4 * (size_t)i;
            rec[4 * (size_t)i + 0] = make_float4(m[0], m[4], m[8], m[12]);
            rec[4 * (size_t)i + 1] = make_float4(m[1], m[5], m[9], m[13]);
            rec[4 * (size_t)i + 2] = make_float4(m[2], m[6], m[10], m[14]);
            int t = s->type[i];
            float tf;
            std::memcpy(&tf, &t, sizeof(tf));
            rec[4 * (size_t)i + 3] = make_float4((float)q[0], (float)q[1], (float)q[2], tf);
        }
    }
    if ((rc = upload(c, rec.data(), rec.size(), &ds.rec))) return rc;

    // acceleration structure: host-supplied or built natively
    std::vector<rm_bvh_node> bvh;
    std::vector<rm_octree_node> oct;
    std::vector<int32_t> leaf;
    if (s->accel_kind != RM_ACCEL_NONE) {
        if (s->nodes) {
            if (s->n_nodes <= 0 || s->n_leaf_prims < 0 || (s->n_leaf_prims > 0 && !s->leaf_prim_index))
                return fail(c, RM_ERR_ARG, "bad acceleration-structure arrays");
            leaf.assign(s->leaf_prim_index, s->leaf_prim_index + s->n_leaf_prims);
            if (s->accel_kind == RM_ACCEL_BVH) bvh.assign((const rm_bvh_node*)s->nodes, (const rm_bvh_node*)s->nodes + s->n_nodes);
            else oct.assign((const rm_octree_node*)s->nodes, (const rm_octree_node*)s->nodes + s->n_nodes);
        } else {
            std::vector<PrimGeom> geom;
            compute_object_geometry(*s, c->flags, geom);
            if (s->accel_kind == RM_ACCEL_BVH) build_bvh(geom, bvh, leaf);
            else build_octree(geom, oct, leaf);
        }
        // structural validation (indices in range) so a bad host structure cannot fault the GPU
        const int32_t nn = (int32_t)(s->accel_kind == RM_ACCEL_BVH ? bvh.size() : oct.size());
        for (int32_t v : leaf)
            if (v < 0 || v >= nObj) return fail(c, RM_ERR_ARG, "leaf primitive index %d out of range", v);
        if (s->accel_kind == RM_ACCEL_BVH) {
            for (const rm_bvh_node& nd : bvh) {
                if (nd.left >= nn || nd.right >= nn || nd.left < -1 || nd.right < -1) return fail(c, RM_ERR_ARG, "BVH child index out of range");
                if (nd.prim_count < 0 || nd.prim_first < 0 || (size_t)nd.prim_first + (size_t)nd.prim_count > leaf.size())
                    return fail(c, RM_ERR_ARG, "BVH leaf range out of bounds");
            }
        } else {
            for (const rm_octree_node& nd : oct) {
                if (nd.first_child >= 0 && nd.first_child + 8 > nn) return fail(c, RM_ERR_ARG, "octree child index out of range");
                if (nd.prim_count < 0 || nd.prim_first < 0 || (size_t)nd.prim_first + (size_t)nd.prim_count > leaf.size())
                    return fail(c, RM_ERR_ARG, "octree leaf range out of bounds");
            }
        }
        ds.n_nodes = nn;
        if (s->accel_kind == RM_ACCEL_BVH && !(c->flags & RM_F_VALIDATE_FP64) && !exactOnly) {
            // fast path: locate leaf boxes through a uniform grid instead of descending the tree
            LeafGrid grid;
            build_leaf_grid(bvh, grid);
            for (int k = 0; k < 3; ++k) {
                ds.grid_dims[k] = grid.dims[k];
                ds.grid_origin[k] = grid.origin[k];
                ds.grid_inv[k] = grid.inv_cell[k];
                ds.grid_cell[k] = grid.cell[k];
            }
            ds.lazy_cap = RM_PEND_CAP;
            if (const char* e = std::getenv("RM_LAZY_CAP")) ds.lazy_cap = std::max(1, std::min(RM_PEND_CAP, std::atoi(e)));  // test knob
            std::vector<uint32_t> cellNode(grid.cell_leaf.size());
            for (size_t e = 0; e < cellNode.size(); ++e) cellNode[e] = (uint32_t)grid.leaves[(size_t)grid.cell_leaf[e]].node;
            if ((rc = upload(c, grid.cell_start.data(), grid.cell_start.size(), &ds.grid_cell_start))) return rc;
            if ((rc = upload(c, cellNode.data(), cellNode.size(), &ds.grid_cell_node))) return rc;
            if (allTS) {  // leaf records of the point query
                std::vector<LeafRecTS> lrec(bvh.size());
                std::memset(lrec.data(), 0, lrec.size() * sizeof(LeafRecTS));
                for (size_t ni = 0; ni < bvh.size(); ++ni) {
                    const rm_bvh_node& nd = bvh[ni];
                    if (nd.left >= 0 || nd.right >= 0 || nd.prim_count <= 0) continue;
                    LeafRecTS& lr = lrec[ni];
                    for (int k = 0; k < 3; ++k) {
                        lr.bmin[k] = nd.bmin[k];
                        lr.bmax[k] = nd.bmax[k];
                    }
                    lr.prim_first = nd.prim_first;
                    lr.count = nd.prim_count <= 2 ? nd.prim_count : -nd.prim_count;
                    for (int k = 0; k < nd.prim_count && k < 2; ++k) {
                        const int32_t j = leaf[(size_t)nd.prim_first + (size_t)k];
                        const float* m = s->world_to_local + 16 * (size_t)j;
                        lr.s[k] = make_float4(m[12], m[13], m[14], (float)s->params[4 * (size_t)j]);
                        lr.r[k] = s->params[4 * (size_t)j];
                    }
                }
                if ((rc = upload(c, lrec.data(), lrec.size(), &ds.grid_leafrec))) return rc;
                CU(c, cudaStreamSynchronize(c->stream));  // `lrec` goes out of scope
            }
            if (grid.dir_ok) {
                if ((rc = upload(c, reinterpret_cast<const uint4*>(grid.cell_dir.data()), grid.cell_dir.size(), &ds.grid_cell_dir))) return rc;
                if ((rc = upload(c, grid.dir_node.data(), grid.dir_node.size(), &ds.grid_dir_node))) return rc;
            }
            CU(c, cudaStreamSynchronize(c->stream));  // `grid` goes out of scope
        }
        if ((rc = upload(c, bvh.data(), bvh.size(), &ds.bvh))) return rc;
        if ((rc = upload(c, oct.data(), oct.size(), &ds.oct))) return rc;
        if ((rc = upload(c, leaf.data(), leaf.size(), &ds.leaf_prims))) return rc;
    }
    CU(c, cudaStreamSynchronize(c->stream));  // host vectors go out of scope
    c->scene = ds;
    c->exact_only = exactOnly;
    c->has_scene = true;
    return RM_OK;
}

int rm_render_device(rm_ctx* c, const rm_request* rq, const rm_result* out, void* cuda_stream) {
    if (!c) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    return render_device_locked(c, rq, out, cuda_stream ? (cudaStream_t)cuda_stream : c->stream);
}

int rm_render(rm_ctx* c, const rm_request* rq, const rm_result* out) {
    if (!c) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    auto w0 = std::chrono::steady_clock::now();
    int rc = validate_request(c, rq);
    if (rc) return rc;
    if (!out || !out->depth || !out->normal || !out->sdf_eval || !out->iters)
        return fail(c, RM_ERR_ARG, "result planes depth/normal/sdf_eval/iters are required");
    const int bandH = rq->y_end > rq->y_start ? rq->y_end - rq->y_start : 0;
    const size_t np = (size_t)rq->width * (size_t)bandH;
    CU(c, cudaSetDevice(c->device));
    // plane layout inside one staging block (16-byte aligned sections)
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const bool wantRgba = rq->shader != RM_SHADER_NONE && out->rgba, wantRgba2 = rq->shader_analytics != RM_SHADER_NONE && out->rgba_analytics;
    size_t off = 0;
    const size_t oDepth = off; off += al(np);
    const size_t oNormal = off; off += al(3 * np);
    const size_t oSdf = off; off += al(2 * np);
    const size_t oIters = off; off += al(2 * np);
    const size_t oRgba = off; off += wantRgba ? al(4 * np) : 0;
    const size_t oRgba2 = off; off += wantRgba2 ? al(4 * np) : 0;
    const size_t oDf = off; off += out->depth_f32 ? al(4 * np) : 0;
    const size_t oSu = off; off += out->sdf_eval_u32 ? al(4 * np) : 0;
    const size_t oD64 = off; off += out->depth_f64 ? al(8 * np) : 0;
    const size_t total = off;
    if ((rc = ensure(c, c->d_frame, total + 256, false))) return rc;
    uint8_t* d = (uint8_t*)c->d_frame.p;
    rm_result dev{};
    dev.depth = d + oDepth;
    dev.normal = d + oNormal;
    dev.sdf_eval = (uint16_t*)(d + oSdf);
    dev.iters = (uint16_t*)(d + oIters);
    dev.rgba = wantRgba ? d + oRgba : nullptr;
    dev.rgba_analytics = wantRgba2 ? d + oRgba2 : nullptr;
    dev.depth_f32 = out->depth_f32 ? (float*)(d + oDf) : nullptr;
    dev.sdf_eval_u32 = out->sdf_eval_u32 ? (uint32_t*)(d + oSu) : nullptr;
    dev.depth_f64 = out->depth_f64 ? (double*)(d + oD64) : nullptr;
    auto pinned = [&](const void* p, size_t bytes) {
        if (!p) return true;
        for (auto* v : {&c->host_allocs, &c->host_registered, &c->host_shared})
            for (auto& h : *v)
                if ((const char*)p >= h.first && (const char*)p + bytes <= h.first + h.second) return true;
        return false;
    };
    const bool direct = pinned(out->depth, np) && pinned(out->normal, 3 * np) && pinned(out->sdf_eval, 2 * np) && pinned(out->iters, 2 * np) &&
                        pinned(wantRgba ? out->rgba : nullptr, 4 * np) && pinned(wantRgba2 ? out->rgba_analytics : nullptr, 4 * np) &&
                        pinned(out->depth_f32, 4 * np) && pinned(out->sdf_eval_u32, 4 * np) && pinned(out->depth_f64, 8 * np);
    // big frames into page-locked planes: the D2H of finished row bands overlaps the rest of the render
    rm_ctx::EarlyCopy& ec = c->early;
    ec = rm_ctx::EarlyCopy();
    // With row stripes (rq->stripe_count > 1) only the OWNED rows of the band are ever written to the caller's planes, so
    // several contexts (one per GPU, one process each) can fill one shared host frame without touching each other's rows.
    const bool striped = rq->stripe_count > 1;
    ec.band_h = bandH;
    ec.width = rq->width;
    if (striped) {
        ec.stripe_rows = rq->stripe_rows;
        ec.stripe_count = rq->stripe_count;
        ec.stripe_index = rq->stripe_index;
    }
    {
        auto add = [&](void* dst, size_t off, size_t bpp) {
            if (dst) ec.planes[ec.n_planes++] = {(char*)dst, (const char*)d + off, bpp};
        };
        add(out->depth, oDepth, 1);
        add(out->normal, oNormal, 3);
        add(out->sdf_eval, oSdf, 2);
        add(out->iters, oIters, 2);
        add(wantRgba ? out->rgba : nullptr, oRgba, 4);
        add(wantRgba2 ? out->rgba_analytics : nullptr, oRgba2, 4);
        add(out->depth_f32, oDf, 4);
        add(out->sdf_eval_u32, oSu, 4);
        add(out->depth_f64, oD64, 8);
    }
    if (direct && total >= kEarlyCopyMinBytes && early_copy_enabled()) {
        ec.n_bands = (int)std::min<size_t>(kMaxBands, total / (kEarlyCopyMinBytes / 2));
        ec.band_rows = (bandH + ec.n_bands - 1) / ec.n_bands;
        if (striped) {  // whole stripe periods per band: every band then holds the same pattern of owned stripes
            const int period = rq->stripe_rows * rq->stripe_count;
            ec.band_rows = (ec.band_rows + period - 1) / period * period;
        }
        ec.n_bands = (bandH + ec.band_rows - 1) / ec.band_rows;
        ec.on = ec.n_bands > 1;
    }
    rc = render_device_locked(c, rq, &dev, c->stream);
    const bool copied = ec.on;
    ec.on = false;
    if (rc) return rc;
    if (np > 0 && direct && copied) {
        // every band already went out during the render
    } else if (np > 0 && direct) {
        // caller's planes are page-locked (rm_host_alloc / rm_host_register): DMA straight into them
        CU(c, copy_owned_rows(ec, 0, bandH, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
    } else if (np > 0) {
        // one D2H of the whole staging block into pinned memory, then scatter the owned rows to the caller's planes
        if ((rc = ensure(c, c->h_frame, total + 256, true))) return rc;
        CU(c, cudaMemcpyAsync(c->h_frame.p, d, total, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        const char* h = (const char*)c->h_frame.p;
        for (int k = 0; k < ec.n_planes; ++k) {
            const size_t rowB = (size_t)rq->width * ec.planes[k].bpp;
            const char* src = h + (ec.planes[k].src - (const char*)d);
            char* dst = ec.planes[k].dst;
            for_owned_spans(ec.stripe_rows ? ec.stripe_rows : 1, ec.stripe_count, ec.stripe_index, 0, bandH,
                            [&](int a, int rows) { std::memcpy(dst + a * rowB, src + a * rowB, (size_t)rows * rowB); });
        }
    }
    c->last.wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    return RM_OK;
}

int rm_stats(rm_ctx* c, rm_stats_t* out) {
    if (!c || !out) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    *out = c->last;
    return RM_OK;
}

int rm_shade(rm_ctx* c, int32_t shader, uint8_t* rgba, const uint8_t* depth, const uint8_t* normal, const uint16_t* sdf,
             const uint16_t* iters, int32_t width, int32_t height) {
    if (!c) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!rgba || !depth || !normal || !sdf || !iters || width <= 0 || height <= 0) return fail(c, RM_ERR_ARG, "bad shade arguments");
    if (shader < RM_SHADER_NORMAL || shader > RM_SHADER_ITERATION_HEATMAP) return fail(c, RM_ERR_ARG, "bad shader %d", shader);
    CU(c, cudaSetDevice(c->device));
    const size_t np = (size_t)width * height;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t oD = 0, oN = al(np), oS = oN + al(3 * np), oI = oS + al(2 * np), oR = oI + al(2 * np), total = oR + al(4 * np);
    int rc;
    if ((rc = ensure(c, c->d_frame, total, false))) return rc;
    uint8_t* d = (uint8_t*)c->d_frame.p;
    CU(c, cudaMemcpyAsync(d + oD, depth, np, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(d + oN, normal, 3 * np, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(d + oS, sdf, 2 * np, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(d + oI, iters, 2 * np, cudaMemcpyHostToDevice, c->stream));
    ShadeParams sp{};
    sp.shader = shader;
    sp.n_pixels = (int32_t)np;
    sp.depth = d + oD;
    sp.normal = d + oN;
    sp.sdf = (const uint16_t*)(d + oS);
    sp.iters = (const uint16_t*)(d + oI);
    sp.rgba = d + oR;
    int e = (c->flags & RM_F_VALIDATE_FP64) ? launch_shade_val(sp, c->stream) : launch_shade_fast(sp, c->stream);
    if (e != 0) return fail(c, RM_ERR_CUDA, "shade launch: %s", cudaGetErrorString((cudaError_t)e));
    CU(c, cudaMemcpyAsync(rgba, d + oR, 4 * np, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return RM_OK;
}

int rm_probe_fp32_peak(rm_ctx* c, double* tflops) {
    if (!c || !tflops) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(c, cudaSetDevice(c->device));
    int rc;
    if ((rc = ensure(c, c->d_frame, (size_t)c->n_sms * 8 * 256 * sizeof(float), false))) return rc;
    int e = probe_fp32_peak(c->n_sms, c->stream, (float*)c->d_frame.p, tflops);
    if (e != 0) return fail(c, RM_ERR_CUDA, "fp32 probe: %s", cudaGetErrorString((cudaError_t)e));
    return RM_OK;
}

int rm_host_alloc(rm_ctx* c, size_t bytes, void** host_ptr) {
    if (!c || !host_ptr) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(c, cudaSetDevice(c->device));
    void* p = nullptr;
    CU(c, cudaMallocHost(&p, bytes ? bytes : 1));
    c->host_allocs.emplace_back((char*)p, bytes ? bytes : 1);
    *host_ptr = p;
    return RM_OK;
}
int rm_host_free(rm_ctx* c, void* host_ptr) {
    if (!c) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    for (size_t i = 0; i < c->host_allocs.size(); ++i)
        if (c->host_allocs[i].first == (char*)host_ptr) {
            CU(c, cudaSetDevice(c->device));
            CU(c, cudaStreamSynchronize(c->stream));
            CU(c, cudaFreeHost(host_ptr));
            c->host_allocs.erase(c->host_allocs.begin() + (long)i);
            return RM_OK;
        }
    return fail(c, RM_ERR_ARG, "pointer was not allocated by rm_host_alloc on this context");
}

int rm_host_register(rm_ctx* c, void* host_ptr, size_t bytes) {
    if (!c || !host_ptr || !bytes) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable));
    c->host_registered.emplace_back((char*)host_ptr, bytes);
    return RM_OK;
}
int rm_host_unregister(rm_ctx* c, void* host_ptr) {
    if (!c) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    for (size_t i = 0; i < c->host_registered.size(); ++i)
        if (c->host_registered[i].first == (char*)host_ptr) {
            CU(c, cudaSetDevice(c->device));
            CU(c, cudaStreamSynchronize(c->stream));
            CU(c, cudaHostUnregister(host_ptr));
            c->host_registered.erase(c->host_registered.begin() + (long)i);
            return RM_OK;
        }
    return fail(c, RM_ERR_ARG, "pointer was not registered by rm_host_register on this context");
}

int rm_alloc(rm_ctx* c, size_t bytes, void** dev_ptr) {
    if (!c || !dev_ptr) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(c, cudaSetDevice(c->device));
    void* p = nullptr;
    CU(c, cudaMalloc(&p, bytes ? bytes : 1));
    c->user_allocs.push_back(p);
    *dev_ptr = p;
    return RM_OK;
}
int rm_free(rm_ctx* c, void* dev_ptr) {
    if (!c) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    for (size_t i = 0; i < c->user_allocs.size(); ++i)
        if (c->user_allocs[i] == dev_ptr) {
            CU(c, cudaSetDevice(c->device));
            CU(c, cudaFree(dev_ptr));
            c->user_allocs.erase(c->user_allocs.begin() + (long)i);
            return RM_OK;
        }
    return fail(c, RM_ERR_ARG, "pointer was not allocated by rm_alloc on this context");
}
int rm_ipc_export(rm_ctx* c, void* dev_ptr, uint8_t handle[64]) {
    if (!c || !dev_ptr || !handle) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    CU(c, cudaIpcGetMemHandle(&h, dev_ptr));
    std::memcpy(handle, &h, 64);
    return RM_OK;
}
int rm_ipc_open(rm_ctx* c, const uint8_t handle[64], void** dev_ptr) {
    if (!c || !dev_ptr || !handle) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    CU(c, cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return RM_OK;
}
int rm_ipc_close(rm_ctx* c, void* dev_ptr) {
    if (!c || !dev_ptr) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaIpcCloseMemHandle(dev_ptr));
    return RM_OK;
}
int rm_memcpy_d2h(rm_ctx* c, void* host, const void* dev, size_t bytes) {
    if (!c || (!host && bytes) || (!dev && bytes)) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return RM_OK;
}
int rm_memcpy_h2d(rm_ctx* c, void* dev, const void* host, size_t bytes) {
    if (!c || (!host && bytes) || (!dev && bytes)) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return RM_OK;
}

}  // extern "C"
