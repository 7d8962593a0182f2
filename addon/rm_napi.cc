// rm_napi.cc — Node N-API addon over the C ABI of include/rm.h.
//
// Exposes to JavaScript exactly what the reference's worker pool did (src/main.ts:318-321,444-490 + src/workers/raymarchWorker.ts):
//   addon.uploadScene({type: Uint8Array, worldToLocal: Float32Array, params: Float64Array, accel: 'None'|'Octree'|'BVH',
//                      opNodes?: Uint8Array (rm_op_node records, 128 B each), objectRoot?: Int32Array})
//       -> rm_pool_upload_scene     (replaces `new Scene(accel); scene.loadPreset(i)`, raymarchWorker.ts:37-38, once for every GPU)
//   addon.render(job, camera{rot3: Float32Array(9), origin: Float32Array(3)}) : Promise<Result>
//       -> rm_pool_render on a libuv worker thread (the JS event loop never blocks on CUDA).  The library owns every GPU of the
//          box (RM_DEVICES=0,1,... restricts it): a band Job is served from the pool's frame cache, so the <= 4 concurrent band
//          jobs of one frame (main.ts:452-486) cost ONE render across all GPUs.  Result carries the four typed arrays of
//          raymarchWorker.ts:24-31 (ownership moves to JS, like the reference's transfer list :86-91): views of ONE external
//          ArrayBuffer over page-locked memory (rm_pool_host_alloc); blocks return to a free list when JS garbage-collects the
//          buffer.  Where the runtime forbids external buffers (V8 sandbox / recent Electron) they are plain ArrayBuffers.
//          The Result object is held by a napi_ref while the worker thread fills it.
//   addon.stats() -> rm_pool_stats  (diagnostics of main.ts:527-548 over the whole frame of the last band)
//   addon.deviceCount()
// The `Worker` shim that makes main.ts use this unchanged is ts/gpuWorkerShim.ts.  No Node.js exists in the build image: the
// addon is compiled against addon/napi_min.h and EXECUTED under the mock N-API runtime of tests/napi_mock.cc (tests/
// test_gpu_addon.py), which enforces handle scopes; tools/node_harness.mjs is the one-command check on a box with Node >= 18.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#ifdef RM_HAVE_NODE_API_H
#include <node_api.h>
#else
#include "napi_min.h"
#endif
#include "../include/rm.h"

namespace {
rm_pool* g_rm = nullptr;  // every GPU of the box (or RM_DEVICES) behind one object
std::mutex g_rm_mu;

bool ensure_pool(napi_env env) {
    std::lock_guard<std::mutex> lk(g_rm_mu);
    if (g_rm) return true;
    std::vector<int> devs;
    if (const char* e = std::getenv("RM_DEVICES")) {
        for (const char* p = e; *p;) {
            char* end = nullptr;
            const long v = std::strtol(p, &end, 10);
            if (end == p) break;
            devs.push_back((int)v);
            p = (*end == ',') ? end + 1 : end;
        }
    }
    if (rm_pool_create(&g_rm, devs.empty() ? nullptr : devs.data(), (int)devs.size(), 0) != RM_OK) {
        napi_throw_error(env, "RM_ERR_CUDA", rm_pool_last_error(nullptr));
        g_rm = nullptr;
        return false;
    }
    return true;
}

double get_num(napi_env env, napi_value obj, const char* key, double dflt) {
    bool has = false;
    napi_has_named_property(env, obj, key, &has);
    if (!has) return dflt;
    napi_value v;
    double d = dflt;
    if (napi_get_named_property(env, obj, key, &v) != napi_ok || napi_get_value_double(env, v, &d) != napi_ok) return dflt;
    return d;
}
std::string get_str(napi_env env, napi_value obj, const char* key) {
    napi_value v;
    char buf[64] = {0};
    size_t n = 0;
    if (napi_get_named_property(env, obj, key, &v) == napi_ok) napi_get_value_string_utf8(env, v, buf, sizeof(buf), &n);
    return std::string(buf, n);
}
void* get_ta(napi_env env, napi_value obj, const char* key, size_t* len) {
    napi_value v;
    void* data = nullptr;
    napi_typedarray_type t;
    napi_value ab;
    size_t off;
    *len = 0;
    if (napi_get_named_property(env, obj, key, &v) != napi_ok) return nullptr;
    if (napi_get_typedarray_info(env, v, &t, len, &data, &ab, &off) != napi_ok) return nullptr;
    return data;
}
int algorithm_id(const std::string& a) {  // raymarchWorker.ts:49-68 (unknown -> sphere tracer)
    if (a == "fixed-step") return RM_ALG_FIXED_STEP;
    if (a == "adaptive-step") return RM_ALG_ADAPTIVE_STEP;
    if (a == "adaptive-step-v2") return RM_ALG_ADAPTIVE_STEP_V2;
    if (a == "adaptive-step-v3") return RM_ALG_ADAPTIVE_STEP_V3;
    return RM_ALG_SPHERE_TRACER;
}
int accel_id(const std::string& a) { return a == "Octree" ? RM_ACCEL_OCTREE : (a == "BVH" ? RM_ACCEL_BVH : RM_ACCEL_NONE); }

napi_value UploadScene(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value arg;
    napi_get_cb_info(env, info, &argc, &arg, nullptr, nullptr);
    if (!ensure_pool(env)) return nullptr;
    rm_scene s;
    std::memset(&s, 0, sizeof(s));
    size_t nt, nm, np;
    s.type = (const uint8_t*)get_ta(env, arg, "type", &nt);
    s.world_to_local = (const float*)get_ta(env, arg, "worldToLocal", &nm);
    s.params = (const double*)get_ta(env, arg, "params", &np);
    s.n_prims = (int32_t)nt;
    s.accel_kind = accel_id(get_str(env, arg, "accel"));
    if (nm != 16 * nt || np != 4 * nt) {
        napi_throw_error(env, "RM_ERR_ARG", "worldToLocal must hold 16 and params 4 values per primitive");
        return nullptr;
    }
    // operator trees (src/util/primitive_operations): flat rm_op_node records + the root node of every scene object
    size_t nb = 0, nr = 0;
    const void* opn = get_ta(env, arg, "opNodes", &nb);
    const void* roots = get_ta(env, arg, "objectRoot", &nr);
    if (opn && roots && nr > 0) {
        if (nb % sizeof(rm_op_node) != 0) {
            napi_throw_error(env, "RM_ERR_ARG", "opNodes must be a whole number of 128-byte rm_op_node records");
            return nullptr;
        }
        s.op_nodes = (const rm_op_node*)opn;
        s.n_op_nodes = (int32_t)(nb / sizeof(rm_op_node));
        s.object_root = (const int32_t*)roots;
        s.n_objects = (int32_t)nr;
    }
    if (rm_pool_upload_scene(g_rm, &s) != RM_OK) napi_throw_error(env, "RM_ERR", rm_pool_last_error(g_rm));
    return nullptr;
}

// Page-locked result blocks, recycled: cudaMallocHost costs milliseconds, a frame is rendered every few.
std::mutex g_pool_mu;
std::vector<std::pair<void*, size_t>> g_pool;
void* pool_take(size_t bytes, size_t* cap) {
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        for (size_t i = 0; i < g_pool.size(); ++i)
            if (g_pool[i].second >= bytes && g_pool[i].second <= 2 * bytes + 4096) {
                void* p = g_pool[i].first;
                *cap = g_pool[i].second;
                g_pool.erase(g_pool.begin() + (long)i);
                return p;
            }
    }
    void* p = nullptr;
    *cap = bytes;
    return (g_rm && rm_pool_host_alloc(g_rm, bytes, &p) == RM_OK) ? p : nullptr;
}
struct PoolBlock {
    void* p;
    size_t cap;
};
void pool_give(napi_env, void*, void* hint) {  // napi_finalize of the external ArrayBuffer (JS thread)
    PoolBlock* b = (PoolBlock*)hint;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_pool.emplace_back(b->p, b->cap);
    delete b;
}

struct RenderWork {
    rm_request rq;
    rm_result out;
    napi_deferred deferred;
    napi_async_work work;
    // The Result object (and through it the ArrayBuffer the worker thread writes into) must outlive the native call that
    // created it: a napi_value dies with its handle scope, so a strong reference carries it to the completion callback.
    napi_ref keep;
    int rc;
    std::string err;
    size_t npx;
    // ArrayBuffers are created on the JS thread before the work is queued; the worker thread only fills them
    void *depth, *normal, *sdf, *iters;
};

void RenderExecute(napi_env, void* data) {  // libuv worker thread: no JS here
    RenderWork* w = (RenderWork*)data;
    w->rc = rm_pool_render(g_rm, &w->rq, &w->out);
    if (w->rc != RM_OK) w->err = rm_pool_last_error(g_rm);
}
void RenderComplete(napi_env env, napi_status, void* data) {  // JS thread, inside a fresh handle scope
    RenderWork* w = (RenderWork*)data;
    napi_value result_obj = nullptr;
    const bool have = napi_get_reference_value(env, w->keep, &result_obj) == napi_ok && result_obj != nullptr;
    if (w->rc == RM_OK && have) {
        napi_resolve_deferred(env, w->deferred, result_obj);
    } else {
        if (w->rc == RM_OK) w->err = "the Result object was lost";
        napi_value msg, e;
        napi_create_string_utf8(env, w->err.c_str(), w->err.size(), &msg);
        napi_create_error(env, nullptr, msg, &e);
        napi_reject_deferred(env, w->deferred, e);
    }
    napi_delete_reference(env, w->keep);
    napi_delete_async_work(env, w->work);
    delete w;
}

napi_value Render(napi_env env, napi_callback_info info) {
    size_t argc = 2;
    napi_value argv[2];
    napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
    napi_value job = argv[0], cam = argv[1];
    if (!ensure_pool(env)) return nullptr;
    RenderWork* w = new RenderWork();
    std::memset(&w->rq, 0, sizeof(w->rq));
    std::memset(&w->out, 0, sizeof(w->out));
    w->rq.width = (int32_t)get_num(env, job, "width", 0);
    w->rq.height = (int32_t)get_num(env, job, "height", 0);
    w->rq.y_start = (int32_t)get_num(env, job, "yStart", 0);
    w->rq.y_end = (int32_t)get_num(env, job, "yEnd", w->rq.height);
    w->rq.time = get_num(env, job, "time", 0);
    w->rq.algorithm = algorithm_id(get_str(env, job, "algorithm"));
    w->rq.step_size = get_num(env, job, "stepSize", 0.1);
    w->rq.overshoot_factor = get_num(env, job, "overshootFactor", 1.2);
    w->rq.shader = RM_SHADER_NONE;  // main.ts shades on its own thread (main.ts:493-515); fuse by passing a shader id
    w->rq.shader_analytics = RM_SHADER_NONE;
    size_t n9, n3;
    const float* rot3 = (const float*)get_ta(env, cam, "rot3", &n9);
    const float* org = (const float*)get_ta(env, cam, "origin", &n3);
    if (!rot3 || !org || n9 != 9 || n3 != 3) {
        delete w;
        napi_throw_error(env, "RM_ERR_ARG", "camera.rot3 (Float32Array 9) and camera.origin (Float32Array 3) required");
        return nullptr;
    }
    std::memcpy(w->rq.rot3, rot3, sizeof(w->rq.rot3));
    std::memcpy(w->rq.origin, org, sizeof(w->rq.origin));
    const int th = w->rq.y_end > w->rq.y_start ? w->rq.y_end - w->rq.y_start : 0;
    w->npx = (size_t)w->rq.width * th;
    // Result arrays (raymarchWorker.ts:42-46): four views of one ArrayBuffer, sections 256-byte aligned.  Preferred backing:
    // a page-locked block (direct DMA + early band download); otherwise a plain ArrayBuffer (rm_render stages the copy).
    napi_value ab = nullptr, ta = nullptr, res = nullptr, v = nullptr;
    napi_create_object(env, &res);
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t oDepth = 0, oNormal = al(w->npx), oSdf = oNormal + al(3 * w->npx), oIters = oSdf + al(2 * w->npx);
    const size_t total = oIters + al(2 * w->npx);
    size_t cap = 0;
    void* base = pool_take(total, &cap);
    bool external = false;
    if (base) {
        PoolBlock* blk = new PoolBlock{base, cap};
        external = napi_create_external_arraybuffer(env, base, total, pool_give, blk, &ab) == napi_ok;
        if (!external) {  // napi_no_external_buffers_allowed: keep the block for later, use JS memory
            pool_give(env, nullptr, blk);
            base = nullptr;
        }
    }
    if (!external) napi_create_arraybuffer(env, total, &base, &ab);
    w->depth = (char*)base + oDepth;
    w->normal = (char*)base + oNormal;
    w->sdf = (char*)base + oSdf;
    w->iters = (char*)base + oIters;
    napi_create_typedarray(env, napi_uint8_clamped_array, w->npx, ab, oDepth, &ta);
    napi_set_named_property(env, res, "depth", ta);
    napi_create_typedarray(env, napi_uint8_clamped_array, 3 * w->npx, ab, oNormal, &ta);
    napi_set_named_property(env, res, "normal", ta);
    napi_create_typedarray(env, napi_uint16_array, w->npx, ab, oSdf, &ta);
    napi_set_named_property(env, res, "sdfEval", ta);
    napi_create_typedarray(env, napi_uint16_array, w->npx, ab, oIters, &ta);
    napi_set_named_property(env, res, "iters", ta);
    napi_create_int32(env, w->rq.y_start, &v);
    napi_set_named_property(env, res, "yStart", v);
    napi_create_int32(env, w->rq.y_end, &v);
    napi_set_named_property(env, res, "yEnd", v);
    w->out.depth = (uint8_t*)w->depth;
    w->out.normal = (uint8_t*)w->normal;
    w->out.sdf_eval = (uint16_t*)w->sdf;
    w->out.iters = (uint16_t*)w->iters;
    if (napi_create_reference(env, res, 1, &w->keep) != napi_ok) {
        delete w;
        napi_throw_error(env, "RM_ERR", "napi_create_reference failed");
        return nullptr;
    }
    napi_value promise, name;
    napi_create_promise(env, &w->deferred, &promise);
    napi_create_string_utf8(env, "rm_render", 9, &name);
    napi_create_async_work(env, nullptr, name, RenderExecute, RenderComplete, w, &w->work);
    napi_queue_async_work(env, w->work);
    return promise;
}

napi_value Stats(napi_env env, napi_callback_info) {
    rm_stats_t st;
    napi_value o, v;
    napi_create_object(env, &o);
    if (!g_rm || rm_pool_stats(g_rm, &st) != RM_OK) return o;
    napi_create_double(env, (double)st.sum_sdf, &v);
    napi_set_named_property(env, o, "totalSDFCalls", v);
    napi_create_double(env, (double)st.max_sdf, &v);
    napi_set_named_property(env, o, "maxSDFCalls", v);
    napi_create_double(env, (double)st.min_sdf, &v);
    napi_set_named_property(env, o, "minSDFCalls", v);
    napi_create_double(env, (double)st.sum_iters, &v);
    napi_set_named_property(env, o, "totalIterations", v);
    napi_create_double(env, st.kernel_ms, &v);
    napi_set_named_property(env, o, "kernelMs", v);
    napi_create_double(env, (double)st.n_pixels, &v);
    napi_set_named_property(env, o, "totalPixels", v);
    napi_create_int32(env, st.n_devices, &v);
    napi_set_named_property(env, o, "devices", v);
    return o;
}

napi_value DeviceCount(napi_env env, napi_callback_info) {
    napi_value v;
    napi_create_int32(env, ensure_pool(env) ? rm_pool_device_count(g_rm) : 0, &v);
    return v;
}

napi_value Init(napi_env env, napi_value exports) {
    napi_property_descriptor d[] = {{"uploadScene", nullptr, UploadScene, nullptr, nullptr, nullptr, 0, nullptr},
                                    {"render", nullptr, Render, nullptr, nullptr, nullptr, 0, nullptr},
                                    {"stats", nullptr, Stats, nullptr, nullptr, nullptr, 0, nullptr},
                                    {"deviceCount", nullptr, DeviceCount, nullptr, nullptr, nullptr, 0, nullptr}};
    napi_define_properties(env, exports, 4, d);
    return exports;
}
}  // namespace

// Module entry: the well-known symbol Node resolves when it dlopens the .node file (what NAPI_MODULE_INIT() expands to).
extern "C" __attribute__((visibility("default"))) napi_value napi_register_module_v1(napi_env env, napi_value exports) { return Init(env, exports); }
