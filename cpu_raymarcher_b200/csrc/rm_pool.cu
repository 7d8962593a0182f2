// rm_pool.cu — every GPU of the box behind ONE object in ONE process (include/rm.h "rm_pool").
//
// What it replaces: the reference's worker pool — `new Worker(...)` x NUM_WORKERS (src/main.ts:318-321), the per-frame band
// partition + Promise.all (main.ts:444-490), the tile copy into the frame buffers (main.ts:461-468) and the diagnostics loop
// over the assembled frame (main.ts:527-548).  One rm_ctx + one persistent host thread per device; the public rm_* entry
// points do the per-device work, this file only deals the work out and puts the results together.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "rm_ctx.h"

using namespace rm;

namespace {
thread_local std::string g_pool_create_error;
constexpr int kPoolStripeRows = 8;  // rows per interleaved stripe (a multiple of the 4-row tile); 4-row stripes measured 2 % slower (profiles/r02k_stripe_time.log)

// One persistent host thread per device: CUDA calls for device i are always issued from thread i.
struct DeviceThread {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = false, quit = false;
    int rc = 0;
};

// Key of the frame cache: everything of a request except the band.
struct FrameKey {
    int32_t width = 0, height = 0, algorithm = 0, shader = 0, shader_analytics = 0;
    double time = 0, step_size = 0, overshoot = 0;
    float rot3[9] = {}, origin[3] = {};
    unsigned long long scene_gen = 0;
    bool operator==(const FrameKey& o) const {
        return width == o.width && height == o.height && algorithm == o.algorithm && shader == o.shader && shader_analytics == o.shader_analytics &&
               std::memcmp(&time, &o.time, sizeof(double)) == 0 && std::memcmp(&step_size, &o.step_size, sizeof(double)) == 0 &&
               std::memcmp(&overshoot, &o.overshoot, sizeof(double)) == 0 && std::memcmp(rot3, o.rot3, sizeof(rot3)) == 0 &&
               std::memcmp(origin, o.origin, sizeof(origin)) == 0 && scene_gen == o.scene_gen;
    }
};
}  // namespace

struct rm_pool {
    std::vector<rm_ctx*> ctx;
    std::vector<DeviceThread*> thr;
    unsigned flags = 0;
    std::mutex mu;  // one pool call at a time (concurrent band requests queue up here and hit the frame cache)
    std::string err;
    unsigned long long scene_gen = 0;
    bool has_scene = false;
    bool peer_ok = true;  // every device can store into device 0's memory (rm_pool_render_device needs it)
    rm_stats_t last{};
    std::vector<rm_stats_t> last_dev;
    // frame cache of band requests: full-frame planes in page-locked host memory owned by the pool
    bool cache_valid = false;
    FrameKey cache_key{};
    char* cache = nullptr;
    size_t cache_cap = 0;
    rm_stats_t cache_stats{};
    std::vector<rm_stats_t> cache_dev;
    std::vector<std::pair<char*, size_t>> host_allocs, host_registered;
    // device scratch planes of rm_pool_render_frames (per device, grown on demand)
    std::vector<void*> scratch;
    std::vector<size_t> scratch_cap;
};

namespace {

int pfail(rm_pool* p, int code, const char* fmt, ...) {
    char buf[640];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (p) p->err = buf;
    else g_pool_create_error = buf;
    return code;
}

void thread_main(DeviceThread* t, int device) {
    cudaSetDevice(device);
    std::unique_lock<std::mutex> lk(t->mu);
    for (;;) {
        t->cv.wait(lk, [&] { return t->has_job || t->quit; });
        if (t->quit) return;
        std::function<int()> job = std::move(t->job);
        t->has_job = false;
        lk.unlock();
        const int rc = job();
        lk.lock();
        t->rc = rc;
        t->done = true;
        t->cv.notify_all();
    }
}

// Run fn(i) on device thread i for every i in [0, n); returns the first non-zero status (and which device).
int run_all(rm_pool* p, int n, const std::function<int(int)>& fn, int* failed = nullptr) {
    for (int i = 0; i < n; ++i) {
        DeviceThread* t = p->thr[(size_t)i];
        std::lock_guard<std::mutex> lk(t->mu);
        t->job = [&fn, i] { return fn(i); };
        t->has_job = true;
        t->done = false;
        t->cv.notify_all();
    }
    int rc = 0;
    for (int i = 0; i < n; ++i) {
        DeviceThread* t = p->thr[(size_t)i];
        std::unique_lock<std::mutex> lk(t->mu);
        t->cv.wait(lk, [&] { return t->done; });
        if (t->rc != 0 && rc == 0) {
            rc = t->rc;
            if (failed) *failed = i;
        }
    }
    return rc;
}

int fail_from_ctx(rm_pool* p, int rc, int dev_idx, const char* what) {
    const char* m = (dev_idx >= 0 && dev_idx < (int)p->ctx.size()) ? rm_last_error(p->ctx[(size_t)dev_idx]) : "";
    return pfail(p, rc, "%s (device %d): %s", what, dev_idx >= 0 ? p->ctx[(size_t)dev_idx]->device : -1, m ? m : "");
}

// main.ts:527-548 over the whole frame from the devices' shares.
void reduce_stats(const std::vector<rm_stats_t>& d, rm_stats_t& out) {
    std::memset(&out, 0, sizeof(out));
    out.min_sdf = 0xffffffffu;
    out.min_iters = 0xffffffffu;
    for (const rm_stats_t& s : d) {
        if (s.n_pixels == 0) continue;  // a device without rows contributes nothing (its minima are the neutral element)
        out.n_pixels += s.n_pixels;
        out.sum_sdf += s.sum_sdf;
        out.sum_iters += s.sum_iters;
        out.sum_sdf_full += s.sum_sdf_full;
        out.sum_iters_full += s.sum_iters_full;
        for (int k = 0; k < 3; ++k) out.evals_by_type[k] += s.evals_by_type[k];
        out.n_hit += s.n_hit;
        out.operator_flops += s.operator_flops;
        out.algorithmic_flops += s.algorithmic_flops;
        out.executed_flops += s.executed_flops;
        out.fp32_pipe_flops += s.fp32_pipe_flops;
        out.tensor_flops += s.tensor_flops;
        out.tc_passes += s.tc_passes;
        out.tc_requests += s.tc_requests;
        out.tc_items += s.tc_items;
        out.n_launches += s.n_launches;
        out.max_sdf = std::max(out.max_sdf, s.max_sdf);
        out.min_sdf = std::min(out.min_sdf, s.min_sdf);
        out.max_iters = std::max(out.max_iters, s.max_iters);
        out.min_iters = std::min(out.min_iters, s.min_iters);
        if (s.kernel_ms >= out.kernel_ms) {  // the devices run concurrently: the slowest one is the frame
            out.kernel_ms = s.kernel_ms;
            out.drain_ms = s.drain_ms;
            out.tail_ms = s.tail_ms;
        }
    }
    out.device = -1;
    out.n_devices = (int32_t)d.size();
}

size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// Replicate src's resident scene on dst's device with device-to-device peer copies (called on dst's thread).
int clone_scene(rm_ctx* dst, const rm_ctx* src) {
    std::lock_guard<std::mutex> lk(dst->mu);
    cudaError_t e = cudaSetDevice(dst->device);
    if (e != cudaSuccess) {
        dst->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return RM_ERR_CUDA;
    }
    cudaStreamSynchronize(dst->stream);
    free_scene(dst);
    DevScene ds = src->scene;
    for (const SceneAlloc& a : src->scene_allocs) {
        void* q = nullptr;
        if ((e = cudaMalloc(&q, a.bytes)) != cudaSuccess) {
            dst->err = std::string("cudaMalloc (scene replica): ") + cudaGetErrorString(e);
            free_scene(dst);
            return RM_ERR_CUDA;
        }
        dst->scene_allocs.push_back({q, a.bytes, a.field_off});
        if ((e = cudaMemcpyPeerAsync(q, dst->device, a.p, src->device, a.bytes, dst->stream)) != cudaSuccess) {
            dst->err = std::string("cudaMemcpyPeerAsync (scene replica): ") + cudaGetErrorString(e);
            free_scene(dst);
            return RM_ERR_CUDA;
        }
        if (a.field_off >= 0 && (size_t)a.field_off + sizeof(void*) <= sizeof(DevScene)) std::memcpy((char*)&ds + a.field_off, &q, sizeof(void*));
    }
    if ((e = cudaStreamSynchronize(dst->stream)) != cudaSuccess) {
        dst->err = std::string("scene replica copy: ") + cudaGetErrorString(e);
        free_scene(dst);
        return RM_ERR_CUDA;
    }
    dst->scene = ds;
    dst->order_valid = false;
    dst->tree = src->tree;
    dst->d_anim = const_cast<float*>(ds.anim);
    dst->anim_time = src->anim_time;
    dst->anim_valid = src->anim_valid;
    dst->exact_only = src->exact_only;
    dst->has_scene = true;
    return RM_OK;
}

struct PlaneSet {
    size_t off[6];  // depth, normal, sdf, iters, rgba, rgba2
    size_t total;
};
PlaneSet plane_layout(size_t np, bool rgba, bool rgba2) {
    PlaneSet L{};
    size_t o = 0;
    L.off[0] = o; o += al256(np);
    L.off[1] = o; o += al256(3 * np);
    L.off[2] = o; o += al256(2 * np);
    L.off[3] = o; o += al256(2 * np);
    L.off[4] = o; o += rgba ? al256(4 * np) : 0;
    L.off[5] = o; o += rgba2 ? al256(4 * np) : 0;
    L.total = o;
    return L;
}

int check_band(rm_pool* p, const rm_request* rq, const rm_result* out) {
    if (!rq) return pfail(p, RM_ERR_ARG, "request is null");
    if (!out || !out->depth || !out->normal || !out->sdf_eval || !out->iters) return pfail(p, RM_ERR_ARG, "result planes depth/normal/sdf_eval/iters are required");
    if (rq->stripe_count > 1) return pfail(p, RM_ERR_ARG, "rm_pool_render deals the row stripes itself: stripe_count must be 0 or 1");
    if (rq->width <= 0 || rq->height <= 0 || rq->y_start < 0 || rq->y_end > rq->height) return pfail(p, RM_ERR_ARG, "bad frame / band geometry");
    return RM_OK;
}

// All devices render their stripes of the band [rq->y_start, rq->y_end) into the caller's (host or device-0) planes.
int render_striped(rm_pool* p, const rm_request* rq, const rm_result* out, bool device_planes) {
    const int n = (int)p->ctx.size();
    p->last_dev.assign((size_t)n, rm_stats_t{});
    int failed = -1;
    int rc = run_all(p, n, [&](int i) -> int {
        rm_request r = *rq;
        if (n > 1) {
            r.stripe_rows = kPoolStripeRows;
            r.stripe_count = n;
            r.stripe_index = i;
        }
        int e = device_planes ? rm_render_device(p->ctx[(size_t)i], &r, out, nullptr) : rm_render(p->ctx[(size_t)i], &r, out);
        if (e) return e;
        return rm_stats(p->ctx[(size_t)i], &p->last_dev[(size_t)i]);
    }, &failed);
    if (rc) return fail_from_ctx(p, rc, failed, device_planes ? "rm_render_device" : "rm_render");
    reduce_stats(p->last_dev, p->last);
    return RM_OK;
}

}  // namespace

extern "C" {

int rm_pool_create(rm_pool** out, const int* devices, int n_devices, unsigned flags) {
    if (!out) return pfail(nullptr, RM_ERR_ARG, "out is null");
    *out = nullptr;
    const int avail = rm_device_count();
    if (avail <= 0) return pfail(nullptr, RM_ERR_CUDA, "no CUDA device available (%s); librm_b200 has no CPU fallback", rm_last_error(nullptr));
    std::vector<int> devs;
    if (!devices || n_devices <= 0) {
        for (int i = 0; i < avail; ++i) devs.push_back(i);
    } else {
        for (int i = 0; i < n_devices; ++i) {
            if (devices[i] < 0 || devices[i] >= avail) return pfail(nullptr, RM_ERR_ARG, "device %d out of range [0,%d)", devices[i], avail);
            devs.push_back(devices[i]);  // a device may be listed more than once: several contexts (and stripe shares) on one GPU
        }
    }
    rm_pool* p = new (std::nothrow) rm_pool();
    if (!p) return pfail(nullptr, RM_ERR_NOMEM, "out of memory");
    p->flags = flags;
    for (int d : devs) {
        rm_ctx* c = nullptr;
        int rc = rm_create(&c, d, flags);
        if (rc) {
            pfail(nullptr, rc, "rm_create(device %d): %s", d, rm_last_error(nullptr));
            rm_pool_destroy(p);
            return rc;
        }
        p->ctx.push_back(c);
    }
    // peer access both ways between device 0 and the others: scene replication and the fused gather into device 0's planes
    for (size_t i = 1; i < devs.size(); ++i) {
        if (devs[i] == devs[0]) continue;  // same GPU: same address space
        int can = 0;
        cudaDeviceCanAccessPeer(&can, devs[i], devs[0]);
        if (!can) p->peer_ok = false;
        if (can) {
            cudaSetDevice(devs[i]);
            cudaError_t e = cudaDeviceEnablePeerAccess(devs[0], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            cudaSetDevice(devs[0]);
            e = cudaDeviceEnablePeerAccess(devs[i], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        }
        cudaGetLastError();
    }
    for (size_t i = 0; i < devs.size(); ++i) {
        DeviceThread* t = new DeviceThread();
        t->th = std::thread(thread_main, t, devs[i]);
        p->thr.push_back(t);
    }
    p->scratch.assign(devs.size(), nullptr);
    p->scratch_cap.assign(devs.size(), 0);
    p->last_dev.assign(devs.size(), rm_stats_t{});
    *out = p;
    return RM_OK;
}

void rm_pool_destroy(rm_pool* p) {
    if (!p) return;
    for (DeviceThread* t : p->thr) {
        {
            std::lock_guard<std::mutex> lk(t->mu);
            t->quit = true;
            t->cv.notify_all();
        }
        if (t->th.joinable()) t->th.join();
        delete t;
    }
    // (the per-device scratch planes were allocated with rm_alloc: their contexts free them in rm_destroy below)
    if (!p->ctx.empty()) cudaSetDevice(p->ctx[0]->device);
    if (p->cache) cudaFreeHost(p->cache);
    for (auto& h : p->host_allocs) cudaFreeHost(h.first);
    for (auto& h : p->host_registered) cudaHostUnregister(h.first);
    for (rm_ctx* c : p->ctx) rm_destroy(c);
    delete p;
}

const char* rm_pool_last_error(rm_pool* p) { return p ? p->err.c_str() : g_pool_create_error.c_str(); }

int rm_pool_device_count(rm_pool* p) { return p ? (int)p->ctx.size() : RM_ERR_ARG; }

int rm_pool_upload_scene(rm_pool* p, const rm_scene* scene) {
    if (!p) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(p->mu);
    p->has_scene = false;
    p->cache_valid = false;
    ++p->scene_gen;
    const int n = (int)p->ctx.size();
    // device 0: compile operator trees, build BVH / octree / leaf grid / cluster data, upload — once
    int rc = run_all(p, 1, [&](int) { return rm_upload_scene(p->ctx[0], scene); });
    if (rc) return fail_from_ctx(p, rc, 0, "rm_upload_scene");
    // devices 1..n-1: device-to-device replicas over NVLink, concurrently
    if (n > 1) {
        int failed = -1;
        rc = run_all(p, n, [&](int i) { return i == 0 ? RM_OK : clone_scene(p->ctx[(size_t)i], p->ctx[0]); }, &failed);
        if (rc) return fail_from_ctx(p, rc, failed, "scene replica");
    }
    p->has_scene = true;
    return RM_OK;
}

int rm_pool_render_device(rm_pool* p, const rm_request* rq, const rm_result* out) {
    if (!p) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(p->mu);
    if (!p->has_scene) return pfail(p, RM_ERR_STATE, "rm_pool_render_device before rm_pool_upload_scene");
    if (p->ctx.size() > 1 && !p->peer_ok) return pfail(p, RM_ERR_STATE, "the devices of this pool cannot access device 0's memory (no peer access): use rm_pool_render");
    int rc = check_band(p, rq, out);
    if (rc) return rc;
    auto w0 = std::chrono::steady_clock::now();
    rc = render_striped(p, rq, out, true);
    p->last.wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    return rc;
}

int rm_pool_render(rm_pool* p, const rm_request* rq, const rm_result* out) {
    if (!p) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(p->mu);
    if (!p->has_scene) return pfail(p, RM_ERR_STATE, "rm_pool_render before rm_pool_upload_scene");
    int rc = check_band(p, rq, out);
    if (rc) return rc;
    auto w0 = std::chrono::steady_clock::now();
    const int bandH = rq->y_end > rq->y_start ? rq->y_end - rq->y_start : 0;
    const bool isBand = bandH < rq->height;
    const bool extras = out->depth_f32 || out->sdf_eval_u32 || out->depth_f64;
    if (!isBand || extras || bandH == 0) {
        // a whole frame (or a band that wants the extension planes): straight into the caller's planes
        rc = render_striped(p, rq, out, false);
        p->last.wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
        return rc;
    }
    // A band Job of a frame (main.ts:452-486 posts <= 4 of them per frame): the first one of a new key renders the WHOLE frame
    // across all devices into the pool's page-locked frame cache — one launch per device and one end-of-frame tail instead of
    // one per band — and every band request copies its rows out of it.
    FrameKey key{};
    key.width = rq->width;
    key.height = rq->height;
    key.algorithm = rq->algorithm;
    key.shader = rq->shader;
    key.shader_analytics = rq->shader_analytics;
    key.time = rq->time;
    key.step_size = rq->step_size;
    key.overshoot = rq->overshoot_factor;
    std::memcpy(key.rot3, rq->rot3, sizeof(key.rot3));
    std::memcpy(key.origin, rq->origin, sizeof(key.origin));
    key.scene_gen = p->scene_gen;
    const size_t np = (size_t)rq->width * (size_t)rq->height;
    const bool wantRgba = rq->shader != RM_SHADER_NONE, wantRgba2 = rq->shader_analytics != RM_SHADER_NONE;
    const PlaneSet L = plane_layout(np, wantRgba, wantRgba2);
    if (!(p->cache_valid && p->cache_key == key)) {
        p->cache_valid = false;
        if (p->cache_cap < L.total) {
            cudaSetDevice(p->ctx[0]->device);
            if (p->cache) {
                for (rm_ctx* c : p->ctx) c->host_shared.clear();
                cudaFreeHost(p->cache);
                p->cache = nullptr;
                p->cache_cap = 0;
            }
            void* q = nullptr;
            cudaError_t e = cudaHostAlloc(&q, L.total + L.total / 8, cudaHostAllocPortable);
            if (e != cudaSuccess) return pfail(p, RM_ERR_NOMEM, "frame cache (%zu bytes of page-locked memory): %s", L.total, cudaGetErrorString(e));
            p->cache = (char*)q;
            p->cache_cap = L.total + L.total / 8;
            for (rm_ctx* c : p->ctx) {
                c->host_shared.clear();
                c->host_shared.emplace_back(p->cache, p->cache_cap);
                for (auto& h : p->host_allocs) c->host_shared.push_back(h);
                for (auto& h : p->host_registered) c->host_shared.push_back(h);
            }
        }
        rm_request full = *rq;
        full.y_start = 0;
        full.y_end = rq->height;
        rm_result fr{};
        fr.depth = (uint8_t*)(p->cache + L.off[0]);
        fr.normal = (uint8_t*)(p->cache + L.off[1]);
        fr.sdf_eval = (uint16_t*)(p->cache + L.off[2]);
        fr.iters = (uint16_t*)(p->cache + L.off[3]);
        fr.rgba = wantRgba ? (uint8_t*)(p->cache + L.off[4]) : nullptr;
        fr.rgba_analytics = wantRgba2 ? (uint8_t*)(p->cache + L.off[5]) : nullptr;
        rc = render_striped(p, &full, &fr, false);
        if (rc) return rc;
        p->cache_key = key;
        p->cache_valid = true;
        p->cache_stats = p->last;
        p->cache_dev = p->last_dev;
    } else {
        p->last = p->cache_stats;  // diagnostics of the frame this band belongs to
        p->last_dev = p->cache_dev;
    }
    const size_t r0 = (size_t)rq->y_start * (size_t)rq->width, nb = (size_t)bandH * (size_t)rq->width;
    std::memcpy(out->depth, p->cache + L.off[0] + r0, nb);
    std::memcpy(out->normal, p->cache + L.off[1] + 3 * r0, 3 * nb);
    std::memcpy(out->sdf_eval, p->cache + L.off[2] + 2 * r0, 2 * nb);
    std::memcpy(out->iters, p->cache + L.off[3] + 2 * r0, 2 * nb);
    if (wantRgba && out->rgba) std::memcpy(out->rgba, p->cache + L.off[4] + 4 * r0, 4 * nb);
    if (wantRgba2 && out->rgba_analytics) std::memcpy(out->rgba_analytics, p->cache + L.off[5] + 4 * r0, 4 * nb);
    p->last.wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    return RM_OK;
}

int rm_pool_render_frames(rm_pool* p, const rm_request* rqs, int32_t nreq, const rm_result* host_out, rm_stats_t* stats_out) {
    if (!p) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(p->mu);
    if (!p->has_scene) return pfail(p, RM_ERR_STATE, "rm_pool_render_frames before rm_pool_upload_scene");
    if (nreq < 0 || (nreq > 0 && !rqs)) return pfail(p, RM_ERR_ARG, "bad request array");
    for (int32_t k = 0; k < nreq; ++k)
        if (rqs[k].stripe_count > 1) return pfail(p, RM_ERR_ARG, "request %d: stripe_count must be 0 or 1", k);
    auto w0 = std::chrono::steady_clock::now();
    const int n = (int)p->ctx.size();
    std::vector<rm_stats_t> all((size_t)nreq);
    int failed = -1;
    int rc = run_all(p, n, [&](int i) -> int {
        rm_ctx* c = p->ctx[(size_t)i];
        for (int32_t k = i; k < nreq; k += n) {  // frame k on device k % n
            const rm_request& r = rqs[k];
            int e;
            if (host_out) {
                e = rm_render(c, &r, &host_out[k]);
            } else {
                const int bandH = r.y_end > r.y_start ? r.y_end - r.y_start : 0;
                const size_t np = (size_t)std::max(r.width, 0) * (size_t)bandH;
                const PlaneSet L = plane_layout(np, r.shader != RM_SHADER_NONE, r.shader_analytics != RM_SHADER_NONE);
                if (p->scratch_cap[(size_t)i] < L.total + 256) {
                    if (p->scratch[(size_t)i]) rm_free(c, p->scratch[(size_t)i]);
                    p->scratch[(size_t)i] = nullptr;
                    p->scratch_cap[(size_t)i] = 0;
                    if ((e = rm_alloc(c, L.total + 256, &p->scratch[(size_t)i]))) return e;
                    p->scratch_cap[(size_t)i] = L.total + 256;
                }
                char* d = (char*)p->scratch[(size_t)i];
                rm_result dr{};
                dr.depth = (uint8_t*)(d + L.off[0]);
                dr.normal = (uint8_t*)(d + L.off[1]);
                dr.sdf_eval = (uint16_t*)(d + L.off[2]);
                dr.iters = (uint16_t*)(d + L.off[3]);
                dr.rgba = r.shader != RM_SHADER_NONE ? (uint8_t*)(d + L.off[4]) : nullptr;
                dr.rgba_analytics = r.shader_analytics != RM_SHADER_NONE ? (uint8_t*)(d + L.off[5]) : nullptr;
                e = rm_render_device(c, &r, &dr, nullptr);
            }
            if (e) return e;
            if ((e = rm_stats(c, &all[(size_t)k]))) return e;
        }
        return RM_OK;
    }, &failed);
    if (rc) return fail_from_ctx(p, rc, failed, "rm_pool_render_frames");
    if (stats_out)
        for (int32_t k = 0; k < nreq; ++k) {
            stats_out[k] = all[(size_t)k];
            stats_out[k].n_devices = 1;
        }
    // pool-level summary of the batch: totals over the frames; kernel_ms = the busiest device's sum (the devices run concurrently)
    std::vector<double> devMs((size_t)n, 0.0);
    rm_stats_t sum{};
    reduce_stats(all, sum);
    for (int32_t k = 0; k < nreq; ++k) devMs[(size_t)(k % n)] += all[(size_t)k].kernel_ms;
    sum.kernel_ms = 0.0;
    for (double v : devMs) sum.kernel_ms = std::max(sum.kernel_ms, v);
    sum.n_devices = n;
    sum.wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    p->last = sum;
    return RM_OK;
}

int rm_pool_stats(rm_pool* p, rm_stats_t* out) {
    if (!p || !out) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(p->mu);
    *out = p->last;
    return RM_OK;
}

int rm_pool_device_stats(rm_pool* p, int32_t i, rm_stats_t* out) {
    if (!p || !out) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(p->mu);
    if (i < 0 || i >= (int32_t)p->last_dev.size()) return pfail(p, RM_ERR_ARG, "device index %d out of range", i);
    *out = p->last_dev[(size_t)i];
    return RM_OK;
}

int rm_pool_host_alloc(rm_pool* p, size_t bytes, void** host_ptr) {
    if (!p || !host_ptr) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(p->mu);
    cudaSetDevice(p->ctx[0]->device);
    void* q = nullptr;
    cudaError_t e = cudaHostAlloc(&q, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) return pfail(p, RM_ERR_NOMEM, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    p->host_allocs.emplace_back((char*)q, bytes ? bytes : 1);
    for (rm_ctx* c : p->ctx) c->host_shared.emplace_back((char*)q, bytes ? bytes : 1);
    *host_ptr = q;
    return RM_OK;
}

static int drop_host_range(rm_pool* p, std::vector<std::pair<char*, size_t>>& v, void* host_ptr, bool allocated) {
    for (size_t i = 0; i < v.size(); ++i)
        if (v[i].first == (char*)host_ptr) {
            for (rm_ctx* c : p->ctx) {
                cudaSetDevice(c->device);
                cudaStreamSynchronize(c->stream);
                for (size_t k = 0; k < c->host_shared.size(); ++k)
                    if (c->host_shared[k].first == (char*)host_ptr) {
                        c->host_shared.erase(c->host_shared.begin() + (long)k);
                        break;
                    }
            }
            cudaSetDevice(p->ctx[0]->device);
            cudaError_t e = allocated ? cudaFreeHost(host_ptr) : cudaHostUnregister(host_ptr);
            v.erase(v.begin() + (long)i);
            if (e != cudaSuccess) return pfail(p, RM_ERR_CUDA, "%s: %s", allocated ? "cudaFreeHost" : "cudaHostUnregister", cudaGetErrorString(e));
            return RM_OK;
        }
    return pfail(p, RM_ERR_ARG, "pointer is not known to this pool");
}

int rm_pool_host_free(rm_pool* p, void* host_ptr) {
    if (!p) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(p->mu);
    return drop_host_range(p, p->host_allocs, host_ptr, true);
}

int rm_pool_host_register(rm_pool* p, void* host_ptr, size_t bytes) {
    if (!p || !host_ptr || !bytes) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(p->mu);
    cudaSetDevice(p->ctx[0]->device);
    cudaError_t e = cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) return pfail(p, RM_ERR_CUDA, "cudaHostRegister: %s", cudaGetErrorString(e));
    p->host_registered.emplace_back((char*)host_ptr, bytes);
    for (rm_ctx* c : p->ctx) c->host_shared.emplace_back((char*)host_ptr, bytes);
    return RM_OK;
}

int rm_pool_host_unregister(rm_pool* p, void* host_ptr) {
    if (!p) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(p->mu);
    return drop_host_range(p, p->host_registered, host_ptr, false);
}

int rm_pool_alloc(rm_pool* p, size_t bytes, void** dev_ptr) {
    if (!p || !dev_ptr) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(p->mu);
    int rc = rm_alloc(p->ctx[0], bytes, dev_ptr);
    return rc ? fail_from_ctx(p, rc, 0, "rm_alloc") : RM_OK;
}

int rm_pool_free(rm_pool* p, void* dev_ptr) {
    if (!p) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(p->mu);
    for (rm_ctx* c : p->ctx) {  // no device may still be storing into it
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
    }
    int rc = rm_free(p->ctx[0], dev_ptr);
    return rc ? fail_from_ctx(p, rc, 0, "rm_free") : RM_OK;
}

int rm_pool_memcpy_d2h(rm_pool* p, void* host, const void* dev_ptr, size_t bytes) {
    if (!p) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(p->mu);
    int rc = rm_memcpy_d2h(p->ctx[0], host, dev_ptr, bytes);
    return rc ? fail_from_ctx(p, rc, 0, "rm_memcpy_d2h") : RM_OK;
}

int rm_pool_probe_fp32_peak(rm_pool* p, double* tflops) {
    if (!p || !tflops) return RM_ERR_ARG;
    std::lock_guard<std::mutex> lk(p->mu);
    int rc = rm_probe_fp32_peak(p->ctx[0], tflops);
    return rc ? fail_from_ctx(p, rc, 0, "rm_probe_fp32_peak") : RM_OK;
}

}  // extern "C"
