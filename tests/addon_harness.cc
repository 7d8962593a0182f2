// addon_harness.cc — TEST INFRASTRUCTURE: drives addon/rm_napi.cc through the mock Node-API runtime exactly the way
// ts/gpuWorkerShim.ts drives it from JavaScript: uploadScene(...) once, then the row-band Jobs of a frame posted back to back
// (main.ts:444-486: NUM_WORKERS concurrent jobs, one reply each), promises awaited together, Result arrays consumed.
//
//   addon_harness <scene.bin> <out.bin> <W> <H> <algorithm> <accel> <n_bands>
//
// scene.bin: int32 n_prims, n_op_bytes, n_roots, pad | u8 type[n] | f32 w2l[16 n] | f64 params[4 n] | op-node bytes | i32 roots |
//            f32 rot3[9] | f32 origin[3]            (written by tests/test_gpu_addon.py from the host-side Scene / Camera mirror)
// out.bin  : two frames back to back (the second one re-uses the page-locked blocks the first one's finalizers returned), each
//            u8 depth[WH] | u8 normal[3WH] | u16 sdf[WH] | u16 iters[WH], then f64 stats[6] of the last frame.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "napi_mock.h"

extern "C" napi_value napi_register_module_v1(napi_env env, napi_value exports);
using namespace mock;

static void fail(const std::string& m) {
    std::fprintf(stderr, "addon_harness: %s\n", m.c_str());
    std::exit(1);
}
static Val* prop(Val* o, const char* k) {
    auto it = o->props.find(k);
    if (it == o->props.end()) fail(std::string("missing property ") + k);
    return it->second;
}

int main(int argc, char** argv) {
    if (argc < 8) fail("usage: addon_harness scene.bin out.bin W H algorithm accel n_bands");
    const int W = std::atoi(argv[3]), H = std::atoi(argv[4]), nBands = std::atoi(argv[7]);
    const std::string algorithm = argv[5], accel = argv[6];
    // ---- scene file
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) fail("cannot open scene file");
    int32_t hdr[4];
    if (std::fread(hdr, 4, 4, f) != 4) fail("short scene file");
    const int n = hdr[0], nOpBytes = hdr[1], nRoots = hdr[2];
    std::vector<uint8_t> type((size_t)n), ops((size_t)nOpBytes);
    std::vector<float> w2l((size_t)n * 16), cam(12);
    std::vector<double> params((size_t)n * 4);
    std::vector<int32_t> roots((size_t)nRoots);
    bool ok = std::fread(type.data(), 1, type.size(), f) == type.size() && std::fread(w2l.data(), 4, w2l.size(), f) == w2l.size() &&
              std::fread(params.data(), 8, params.size(), f) == params.size() && std::fread(ops.data(), 1, ops.size(), f) == ops.size() &&
              std::fread(roots.data(), 4, roots.size(), f) == roots.size() && std::fread(cam.data(), 4, 12, f) == 12;
    std::fclose(f);
    if (!ok) fail("short scene file");

    // ---- require('rm_napi.node')
    open_scope();
    Val* exports = make_object();
    napi_register_module_v1(nullptr, wrap(exports));
    open_scope();

    // ---- addon.uploadScene({...})
    Val* scene = make_object();
    scene->props["type"] = make_typedarray(napi_uint8_array, type.data(), type.size(), 1);
    scene->props["worldToLocal"] = make_typedarray(napi_float32_array, w2l.data(), w2l.size(), 4);
    scene->props["params"] = make_typedarray(napi_float64_array, params.data(), params.size(), 8);
    scene->props["accel"] = make_string(accel);
    if (nRoots > 0) {
        scene->props["opNodes"] = make_typedarray(napi_uint8_array, ops.data(), ops.size(), 1);
        scene->props["objectRoot"] = make_typedarray(napi_int32_array, roots.data(), roots.size(), 4);
    }
    call(prop(exports, "uploadScene"), {scene});
    if (!g_exception.empty()) fail("uploadScene threw: " + g_exception);
    Val* dc = call(prop(exports, "deviceCount"), {});
    std::printf("devices %d\n", (int)dc->num);

    Val* camera = make_object();
    camera->props["rot3"] = make_typedarray(napi_float32_array, cam.data(), 9, 4);
    camera->props["origin"] = make_typedarray(napi_float32_array, cam.data() + 9, 3, 4);

    FILE* out = std::fopen(argv[2], "wb");
    if (!out) fail("cannot open output file");
    const size_t np = (size_t)W * H;
    int finalized = 0;
    for (int frame = 0; frame < 2; ++frame) {
        // ---- main.ts:444-486: partition rows over the workers, post every Job, await them all
        const int rowsPer = (H + nBands - 1) / nBands;  // Math.ceil(height / NUM_WORKERS)
        std::vector<Val*> promises;
        std::vector<int> y0s, y1s;
        for (int i = 0; i < nBands; ++i) {
            const int y0 = std::min(i * rowsPer, H), y1 = std::min((i + 1) * rowsPer, H);
            if (y0 >= y1) continue;
            Val* job = make_object();
            job->props["width"] = make_number(W);
            job->props["height"] = make_number(H);
            job->props["time"] = make_number(0);
            job->props["yStart"] = make_number(y0);
            job->props["yEnd"] = make_number(y1);
            job->props["algorithm"] = make_string(algorithm);
            job->props["overshootFactor"] = make_number(1.2);
            job->props["stepSize"] = make_number(0.1);
            Val* p = call(prop(exports, "render"), {job, camera});
            if (!g_exception.empty()) fail("render threw: " + g_exception);
            if (!p || p->kind != Val::Promise) fail("render did not return a promise");
            promises.push_back(p);
            y0s.push_back(y0);
            y1s.push_back(y1);
        }
        drain();  // Promise.all
        std::vector<uint8_t> depth(np), normal(3 * np);
        std::vector<uint16_t> sdf(np), iters(np);
        for (size_t i = 0; i < promises.size(); ++i) {
            Val* p = promises[i];
            if (p->state == 2) fail("render rejected: " + p->result->str);
            if (p->state != 1) fail("promise still pending after the completion callbacks ran");
            Val* r = p->result;
            const size_t nb = (size_t)(y1s[i] - y0s[i]) * W, o = (size_t)y0s[i] * W;
            if ((int)prop(r, "yStart")->num != y0s[i] || (int)prop(r, "yEnd")->num != y1s[i]) fail("Result.yStart / yEnd mismatch");
            auto view = [&](const char* k, int want_type, size_t want_len) -> const char* {
                Val* ta = prop(r, k);
                if (ta->kind != Val::TypedArray || ta->ta_type != want_type || ta->len != want_len) fail(std::string("Result.") + k + " has the wrong type or length");
                return (const char*)ta->ab->data + ta->offset;
            };
            std::memcpy(depth.data() + o, view("depth", napi_uint8_clamped_array, nb), nb);
            std::memcpy(normal.data() + 3 * o, view("normal", napi_uint8_clamped_array, 3 * nb), 3 * nb);
            std::memcpy(sdf.data() + o, view("sdfEval", napi_uint16_array, nb), 2 * nb);
            std::memcpy(iters.data() + o, view("iters", napi_uint16_array, nb), 2 * nb);
        }
        std::fwrite(depth.data(), 1, np, out);
        std::fwrite(normal.data(), 1, 3 * np, out);
        std::fwrite(sdf.data(), 2, np, out);
        std::fwrite(iters.data(), 2, np, out);
        if (g_live_refs != 0) fail("napi_ref leak: " + std::to_string(g_live_refs) + " live references after the frame");
        finalized += gc();  // JS drops the Result arrays: the external ArrayBuffers' finalizers hand their blocks back
    }
    Val* st = call(prop(exports, "stats"), {});
    const double stats[6] = {prop(st, "totalSDFCalls")->num, prop(st, "maxSDFCalls")->num, prop(st, "minSDFCalls")->num,
                             prop(st, "totalIterations")->num, prop(st, "totalPixels")->num, (double)prop(st, "devices")->num};
    std::fwrite(stats, 8, 6, out);
    std::fclose(out);
    std::printf("finalized_external_buffers %d\nOK\n", finalized);
    return 0;
}
