// rm_device.cuh — the raymarch hot path as one persistent sm_100a kernel.
//
// What it replaces (reference file:line):
//   Raymarcher.runRaymarcher / getSceneDistance / getNormal     src/cpu_algorithms/raymarcher.ts:46-135
//   SphereTracer / FixedStep / AdaptiveStep / V2 / V3 .rayMarch   src/cpu_algorithms/*.ts
//   Scene.getDistance                                           src/util/scene.ts:144-190
//   Primitive.sdf + Sphere/Box/Torus.localSdf                   src/util/primitives/*.ts
//   BoundingBox.contains / intersectRay                         src/acceleration_structures/boundingBox.ts:15-21,69-105
//   BVH.findRayIntersections / onRayMarchStep / getPrimitivesAt src/acceleration_structures/bvh.ts:95-240
//   Octree.findNode / intersectRayBox / marchRay                src/acceleration_structures/octree.ts:195-278
//   the four ShadingModel.shade bodies                          src/util/shading_models/*.ts
//   the diagnostics loop                                        src/main.ts:527-548
//
// Design (B200-first, not a translation):
//   * persistent CTAs; each warp pulls 8x4-pixel tiles from an atomic work queue;
//   * every lane owns one ray and runs it as an explicit state machine
//     (STEP -> scene-distance query -> consume -> ... -> 4 normal queries -> finalize); terminated
//     rays are found with __ballot_sync and the warp takes its next tile once the current one has
//     retired (measured best, RM_INIT_LANES); all scene-distance queries of a warp — march steps,
//     normal taps, V3 bridging taps — funnel through ONE code site;
//   * a query is first resolved through the acceleration structure (few primitives, lane-local; the fast BVH
//     kernels use a uniform grid over the leaf boxes + inline leaf records instead of the tree);
//     queries that need every primitive (no acceleration structure, the BVH "empty candidate set"
//     fallback of scene.ts:173, points outside the octree root) run the all-primitives pass: records
//     streamed through shared memory by TMA bulk copies, or — for translation-only spheres behind a
//     BVH — the CTA-cooperative cluster screen on the tcgen05 tensor cores (tc_pass);
//   * SDF operator trees and the Mandelbulb run as compiled per-object programs in the exact kernels;
//   * per-pixel outputs are quantised exactly like the reference's typed arrays and the diagnostics
//     are reduced in the epilogue (warp reduce, one atomic set per warp).
//
// Instantiated twice over the FIELD model (rm_numeric.cuh): NumJS (fp64, bit-exact) and NumFast (fp32).
// Control arithmetic is double / non-fused in both.
#pragma once
#include <type_traits>

#include "rm_numeric.cuh"
#include "rm_types.h"

namespace rm {

constexpr unsigned kFull = 0xffffffffu;
// CTAs (128 threads) per SM the register allocator must leave room for.  Measured on B200: the BVH kernels want
// registers (the 32-wide unrolled search + interval state; 4 CTAs/SM = 128 registers), everything else wants occupancy.
#ifndef RM_INIT_LANES
#define RM_INIT_LANES 32  // lanes that must be free before a warp refills (32 = retire the whole tile first; measured best)
#endif
#ifndef RM_MIN_BLOCKS_BVH
#define RM_MIN_BLOCKS_BVH 4
#endif
#ifndef RM_MIN_BLOCKS_OTHER
#define RM_MIN_BLOCKS_OTHER 8
#endif
#ifndef RM_TC_BATCH
#define RM_TC_BATCH 128  // requests that trigger a tensor-core pass (<= 128)
#endif
#ifndef RM_VEC_EPILOGUE
#define RM_VEC_EPILOGUE 0  // 1: retired pixels are held in registers and written per whole tile with 16-byte vector stores (shuffle transpose).  Parity-green (profiles/r02c_gpu_tests_vec_epilogue.log) but SLOWER on every config (cfg4 +2 %, cfg5 +14 %, profiles/r02c_ab.jsonl): the scalar stores at retirement are fire-and-forget, the tile write sits on the refill path.  Default 0.
#endif
#ifndef RM_INLINE_WALK
#define RM_INLINE_WALK 0  // 1: the grid-walk cell step inlined into the state machine instead of an out-of-line call (A/B switch)
#endif
#ifndef RM_RESIDENT_B
#define RM_RESIDENT_B 1  // cluster-screen B tiles resident in shared memory when they fit the ring (0: always streamed; A/B switch)
#endif
#ifndef RM_SMEM_STATS
#define RM_SMEM_STATS 1  // diagnostics accumulators per warp in shared memory (0: per-lane registers; A/B switch)
#endif
#ifndef RM_TC_STATIC_QUEUE
#define RM_TC_STATIC_QUEUE 1  // tensor-core instance: the cooperative queue is a compile-time fact, the warp-local search body is dropped (211 -> 146 KB of SASS).  With the ray state scalar-replaced: cfg4 32.4 -> 30.0 ms (profiles/r02g_ab.jsonl); before that change the same switch cost +18 % (spills, profiles/r02d_ab.jsonl).  0: A/B switch.
#endif
#ifndef RM_SLAB_PRETEST
#define RM_SLAB_PRETEST 1  // fp32 conservative pre-test in front of the fp64 slab test of the lazy grid walk (0: A/B switch)
#endif
#ifndef RM_TC_WARPS
#define RM_TC_WARPS 16  // warps per CTA of the translation-only-sphere BVH kernel (one CTA per SM; the first 16 run the tensor-core sweeps)
#endif

enum Phase : int {
    PH_IDLE = 0,
    PH_NEW,         // pixel assigned, ray not yet initialised
    PH_STEP,        // at the top of march-loop index i
    PH_WAIT_MARCH,  // scene-distance query pending: march step
    PH_WAIT_V3B,    // query pending: AdaptiveStepV3 bridging tap (adaptiveStepV3.ts:113-115)
    PH_WAIT_N0,     // query pending: normal taps (raymarcher.ts:124-132)
    PH_WAIT_N1,
    PH_WAIT_N2,
    PH_WAIT_N3,
    PH_FINAL,  // march + normal done: quantise and store
    PH_HELD    // vectorised epilogue: retired, quantised outputs held in registers until the warp's next refill point
};

// ------------------------------------------------------------------------------------------
// Primitive SDFs
// ------------------------------------------------------------------------------------------

// Exact model: Primitive.sdf (primitive.ts:33-39) = f32(transformMat4(p, M)) then localSdf, for leaf primitive j.
RM_DEV double leaf_sdf_exact(const DevScene& sc, int j, double x, double y, double z, int length_sqrt, double time_scale = 1.0) {
    const float* m = sc.w2l + 16 * (size_t)j;
    double m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3], m4 = m[4], m5 = m[5], m6 = m[6], m7 = m[7];
    double m8 = m[8], m9 = m[9], m10 = m[10], m11 = m[11], m12 = m[12], m13 = m[13], m14 = m[14], m15 = m[15];
    double w = m3 * x + m7 * y + m11 * z + m15;
    if (w == 0.0 || w != w) w = 1.0;  // w = w || 1.0
    double lx = (double)f32r((m0 * x + m4 * y + m8 * z + m12) / w);
    double ly = (double)f32r((m1 * x + m5 * y + m9 * z + m13) / w);
    double lz = (double)f32r((m2 * x + m6 * y + m10 * z + m14) / w);
    const double* prm = sc.params + 4 * (size_t)j;
    int type = sc.type[j];
    if (type == RM_PRIM_SPHERE) {  // sphere.ts:12-14 ; vec3.length = Math.hypot (or plain sqrt, RM_F_LENGTH_SQRT)
        double len = length_sqrt ? sqrt(lx * lx + ly * ly + lz * lz) : v8_hypot3(lx, ly, lz);
        return len - prm[0];
    } else if (type == RM_PRIM_BOX) {  // box.ts:13-30
        double q0 = (double)f32r(fabs(lx) - prm[0]);
        double q1 = (double)f32r(fabs(ly) - prm[1]);
        double q2 = (double)f32r(fabs(lz) - prm[2]);
        double o0 = (double)f32r(jsmax(q0, 0.0));
        double o1 = (double)f32r(jsmax(q1, 0.0));
        double o2 = (double)f32r(jsmax(q2, 0.0));
        double outsideDist = length_sqrt ? sqrt(o0 * o0 + o1 * o1 + o2 * o2) : v8_hypot3(o0, o1, o2);
        double insideDist = jsmin(jsmax(q0, jsmax(q1, q2)), 0.0);
        return outsideDist + insideDist;
    } else if (type == RM_PRIM_MANDELBULB) {  // mandelbulb.ts:38-78 (Math.* = CUDA's double-precision libm)
        const double px = lx, py = lz, pz = ly;  // p.xyz = p.xzy
        double zx = px, zy = py, zz = pz;
        double dr = 1.0, r = 0.0;
        const double power = prm[0], dphi = (prm[2] != 0.0) ? (sc.time * time_scale) * prm[3] : 0.0;
        const int iterations = (int)prm[1];
        for (int i = 0; i < iterations; ++i) {
            r = length_sqrt ? sqrt(zx * zx + zy * zy + zz * zz) : v8_hypot3(zx, zy, zz);
            if (r > 2.0) break;
            double theta = atan2(zy, zx);
            double phi = asin(zz / r);
            if (prm[2] != 0.0) phi += dphi;
            dr = pow(r, power - 1.0) * dr * power + 1.0;
            r = pow(r, power);  // the loop variable now holds r^power: that is what the return reads after the last pass
            theta = theta * power;
            phi = phi * power;
            double st, ct, sp, cp;
            sincos(theta, &st, &ct);
            sincos(phi, &sp, &cp);
            zx = (double)f32r(r * ct * cp + px);
            zy = (double)f32r(r * st * cp + py);
            zz = (double)f32r(r * sp + pz);
        }
        return 0.5 * log(r) * r / dr;
    } else {  // torus.ts:14-25
        double qx = sqrt(lx * lx + lz * lz) - prm[0];
        double qy = ly;
        return sqrt(qx * qx + qy * qy) - prm[1];
    }
}

// Math.round: ties toward +Infinity, keeps -0 (V8 Float64Round).
RM_DEV double js_round(double x) {
    const double r = ceil(x);
    return (r - 0.5 > x) ? r - 1.0 : r;
}

// Operator trees (src/util/primitive_operations/*.ts): scene object j's compiled program on the point / distance
// stack machine described in rm_types.h.  Everything in JS-number arithmetic (fp64, f32 vector stores, unfused);
// Math.sin / Math.cos are CUDA's double-precision routines (<= 1-2 ulp; the result is consumed through an f32
// store, twist.ts:27-31).  Stack depths are bounded by the RM_MAX_TREE_DEPTH check at upload.
static __device__ __noinline__ double object_sdf_exact(const DevScene& sc, int j, double x, double y, double z, int length_sqrt) {
    double px[kMaxPointStack], py[kMaxPointStack], pz[kMaxPointStack];
    double ds[kMaxDistStack];
    int sp = 0, dp = 0;
    px[0] = x;
    py[0] = y;
    pz[0] = z;
    ds[0] = 10.0;
    const int end = sc.obj_first[j + 1];
    for (int pc = sc.obj_first[j]; pc < end; ++pc) {
        const DevInstr in = sc.instrs[pc];
        const double X = px[sp], Y = py[sp], Z = pz[sp];
        switch (in.op) {
            case I_PRIM: ds[dp++] = leaf_sdf_exact(sc, in.a, X, Y, Z, length_sqrt, in.c); break;  // c: 0 below an AnimatedTranslate
            case I_XFORM: {  // vec3.transformMat4 into a Float32Array
                const float* m = sc.mats + 16 * (size_t)in.a;
                double w = (double)m[3] * X + (double)m[7] * Y + (double)m[11] * Z + (double)m[15];
                if (w == 0.0 || w != w) w = 1.0;
                ++sp;
                px[sp] = (double)f32r(((double)m[0] * X + (double)m[4] * Y + (double)m[8] * Z + (double)m[12]) / w);
                py[sp] = (double)f32r(((double)m[1] * X + (double)m[5] * Y + (double)m[9] * Z + (double)m[13]) / w);
                pz[sp] = (double)f32r(((double)m[2] * X + (double)m[6] * Y + (double)m[10] * Z + (double)m[14]) / w);
                break;
            }
            case I_POP: sp -= in.a; break;
            case I_TWIST: {  // twist.ts:22-33
                const double k = in.c;
                const double c = cos(k * Y), sn = sin(k * Y);
                ++sp;
                px[sp] = (double)f32r(c * X - sn * Z);
                py[sp] = Y;
                pz[sp] = (double)f32r(sn * X + c * Z);
                break;
            }
            case I_REPEAT: {  // repetition.ts:21-26
                const double s0 = (double)in.v[0], s1 = (double)in.v[1], s2 = (double)in.v[2];
                ++sp;
                px[sp] = (double)f32r(X - s0 * js_round(X / s0));
                py[sp] = (double)f32r(Y - s1 * js_round(Y / s1));
                pz[sp] = (double)f32r(Z - s2 * js_round(Z / s2));
                break;
            }
            case I_SUBV: {  // animatedTranslate.ts:42-44
                const float* o = sc.anim + 4 * (size_t)in.a;
                ++sp;
                px[sp] = (double)f32r(X - (double)o[0]);
                py[sp] = (double)f32r(Y - (double)o[1]);
                pz[sp] = (double)f32r(Z - (double)o[2]);
                break;
            }
            case I_SUBC: ds[dp - 1] = ds[dp - 1] - in.c; break;  // round.ts:23
            case I_SUNION: {                                     // smoothUnion.ts:26-34
                const double d1 = ds[dp - 2], d2 = ds[dp - 1];
                const double k = in.c * 4.0;
                const double h = jsmax(k - fabs(d1 - d2), 0.0);
                ds[dp - 2] = jsmin(d1, d2) - h * h * 0.25 / k;
                --dp;
                break;
            }
            default: {  // I_SSUB, smoothSubstraction.ts:25-33
                const double d1 = ds[dp - 2], d2 = ds[dp - 1];
                const double k = in.c * 4.0;
                const double h = jsmax(k - fabs(d1 + d2), 0.0);
                ds[dp - 2] = jsmax(d1, -d2) + h * h * 0.25 / k;
                --dp;
                break;
            }
        }
    }
    return ds[0];
}

// Scene object j (Scene.objectSDFs[j].sdf(p)): a primitive, or an operator tree over primitives.
RM_DEV double prim_sdf_exact(const DevScene& sc, int j, double x, double y, double z, int length_sqrt) {
    if (sc.n_instrs > 0) return object_sdf_exact(sc, j, x, y, z, length_sqrt);
    return leaf_sdf_exact(sc, j, x, y, z, length_sqrt);
}

// ---- fp32 evaluators of scene objects (operator trees, Mandelbulb) for default contexts (round 2) --------------------------
// The exact kernels serve these scenes in every context; without RM_F_VALIDATE_FP64 the MARCH-STEP queries evaluate the object
// programs in fp32 (same stack machine, float arithmetic, CUDA's single-precision libm) while the four normal taps — whose
// differences of nearly equal distances set the normal bytes — and any query that lands within 2e-5 of the hit threshold are
// evaluated exactly.  Bar: >= 99.9 % of pixels agree (hit mask, RGB within 1/255, depth rel. err <= 1e-4).
RM_DEV float leaf_sdf_f32(const DevScene& sc, int j, float x, float y, float z, float time_scale) {
    const float* m = sc.w2l + 16 * (size_t)j;
    float w = fmaf(m[3], x, fmaf(m[7], y, fmaf(m[11], z, m[15])));
    if (w == 0.f || w != w) w = 1.f;
    const float iw = 1.f / w;
    const float lx = fmaf(m[0], x, fmaf(m[4], y, fmaf(m[8], z, m[12]))) * iw;
    const float ly = fmaf(m[1], x, fmaf(m[5], y, fmaf(m[9], z, m[13]))) * iw;
    const float lz = fmaf(m[2], x, fmaf(m[6], y, fmaf(m[10], z, m[14]))) * iw;
    const double* prm = sc.params + 4 * (size_t)j;
    const int type = sc.type[j];
    if (type == RM_PRIM_SPHERE) {
        return sqrtf(fmaf(lx, lx, fmaf(ly, ly, lz * lz))) - (float)prm[0];
    } else if (type == RM_PRIM_BOX) {
        const float q0 = fabsf(lx) - (float)prm[0], q1 = fabsf(ly) - (float)prm[1], q2 = fabsf(lz) - (float)prm[2];
        const float o0 = fmaxf(q0, 0.f), o1 = fmaxf(q1, 0.f), o2 = fmaxf(q2, 0.f);
        return sqrtf(fmaf(o0, o0, fmaf(o1, o1, o2 * o2))) + fminf(fmaxf(q0, fmaxf(q1, q2)), 0.f);
    } else if (type == RM_PRIM_MANDELBULB) {  // mandelbulb.ts:38-78 in float
        const float px = lx, py = lz, pz = ly;
        float zx = px, zy = py, zz = pz, dr = 1.f, r = 0.f;
        const float power = (float)prm[0], dphi = (prm[2] != 0.0) ? (float)((sc.time * (double)time_scale) * prm[3]) : 0.f;
        const int iterations = (int)prm[1];
        for (int i = 0; i < iterations; ++i) {
            r = sqrtf(fmaf(zx, zx, fmaf(zy, zy, zz * zz)));
            if (r > 2.f) break;
            float theta = atan2f(zy, zx), phi = asinf(zz / r) + dphi;
            const float rp1 = powf(r, power - 1.f);
            dr = fmaf(rp1 * dr, power, 1.f);
            r = rp1 * r;  // r^power: what the return reads after the last pass
            theta *= power;
            phi *= power;
            float st, ct, sp, cp;
            sincosf(theta, &st, &ct);
            sincosf(phi, &sp, &cp);
            zx = fmaf(r * ct, cp, px);
            zy = fmaf(r * st, cp, py);
            zz = fmaf(r, sp, pz);
        }
        return 0.5f * logf(r) * r / dr;
    } else {
        const float qx = sqrtf(fmaf(lx, lx, lz * lz)) - (float)prm[0];
        return sqrtf(fmaf(qx, qx, ly * ly)) - (float)prm[1];
    }
}

static __device__ __noinline__ float object_sdf_f32(const DevScene& sc, int j, float x, float y, float z) {
    float px[kMaxPointStack], py[kMaxPointStack], pz[kMaxPointStack];
    float ds[kMaxDistStack];
    int sp = 0, dp = 0;
    px[0] = x;
    py[0] = y;
    pz[0] = z;
    ds[0] = 10.f;
    const int end = sc.obj_first[j + 1];
    for (int pc = sc.obj_first[j]; pc < end; ++pc) {
        const DevInstr in = sc.instrs[pc];
        const float X = px[sp], Y = py[sp], Z = pz[sp];
        switch (in.op) {
            case I_PRIM: ds[dp++] = leaf_sdf_f32(sc, in.a, X, Y, Z, (float)in.c); break;
            case I_XFORM: {
                const float* m = sc.mats + 16 * (size_t)in.a;
                float w = fmaf(m[3], X, fmaf(m[7], Y, fmaf(m[11], Z, m[15])));
                if (w == 0.f || w != w) w = 1.f;
                const float iw = 1.f / w;
                ++sp;
                px[sp] = fmaf(m[0], X, fmaf(m[4], Y, fmaf(m[8], Z, m[12]))) * iw;
                py[sp] = fmaf(m[1], X, fmaf(m[5], Y, fmaf(m[9], Z, m[13]))) * iw;
                pz[sp] = fmaf(m[2], X, fmaf(m[6], Y, fmaf(m[10], Z, m[14]))) * iw;
                break;
            }
            case I_POP: sp -= in.a; break;
            case I_TWIST: {
                float sn, c;
                sincosf((float)in.c * Y, &sn, &c);
                ++sp;
                px[sp] = c * X - sn * Z;
                py[sp] = Y;
                pz[sp] = sn * X + c * Z;
                break;
            }
            case I_REPEAT: {  // Math.round: ties toward +Infinity
                const float s0 = in.v[0], s1 = in.v[1], s2 = in.v[2];
                ++sp;
                px[sp] = X - s0 * floorf(X / s0 + 0.5f);
                py[sp] = Y - s1 * floorf(Y / s1 + 0.5f);
                pz[sp] = Z - s2 * floorf(Z / s2 + 0.5f);
                break;
            }
            case I_SUBV: {
                const float* o = sc.anim + 4 * (size_t)in.a;
                ++sp;
                px[sp] = X - o[0];
                py[sp] = Y - o[1];
                pz[sp] = Z - o[2];
                break;
            }
            case I_SUBC: ds[dp - 1] = ds[dp - 1] - (float)in.c; break;
            case I_SUNION: {
                const float d1 = ds[dp - 2], d2 = ds[dp - 1], k = (float)in.c * 4.f;
                const float h = fmaxf(k - fabsf(d1 - d2), 0.f);
                ds[dp - 2] = fminf(d1, d2) - h * h * 0.25f / k;
                --dp;
                break;
            }
            default: {  // I_SSUB
                const float d1 = ds[dp - 2], d2 = ds[dp - 1], k = (float)in.c * 4.f;
                const float h = fmaxf(k - fabsf(d1 + d2), 0.f);
                ds[dp - 2] = fmaxf(d1, -d2) + h * h * 0.25f / k;
                --dp;
                break;
            }
        }
    }
    return ds[0];
}
// Scene object j in fp32: a primitive, or an operator tree over primitives.
RM_DEV float prim_sdf_f32_object(const DevScene& sc, int j, float x, float y, float z) {
    if (sc.n_instrs > 0) return object_sdf_f32(sc, j, x, y, z);
    return leaf_sdf_f32(sc, j, x, y, z, 1.f);
}

// Fast model, general affine record: rows of the 3x4 world->local + (p0,p1,p2,type).
RM_DEV float prim_sdf_fast_general(const float4* __restrict__ rec, int j, float x, float y, float z, int& type) {
    const float4 r0 = __ldg(rec + 4 * (size_t)j + 0);
    const float4 r1 = __ldg(rec + 4 * (size_t)j + 1);
    const float4 r2 = __ldg(rec + 4 * (size_t)j + 2);
    const float4 pr = __ldg(rec + 4 * (size_t)j + 3);
    float lx = fmaf(r0.x, x, fmaf(r0.y, y, fmaf(r0.z, z, r0.w)));
    float ly = fmaf(r1.x, x, fmaf(r1.y, y, fmaf(r1.z, z, r1.w)));
    float lz = fmaf(r2.x, x, fmaf(r2.y, y, fmaf(r2.z, z, r2.w)));
    type = __float_as_int(pr.w);
    if (type == RM_PRIM_SPHERE) {
        return NumFast::sqrt_(fmaf(lx, lx, fmaf(ly, ly, lz * lz))) - pr.x;
    } else if (type == RM_PRIM_BOX) {
        float q0 = fabsf(lx) - pr.x, q1 = fabsf(ly) - pr.y, q2 = fabsf(lz) - pr.z;
        float o0 = fmaxf(q0, 0.f), o1 = fmaxf(q1, 0.f), o2 = fmaxf(q2, 0.f);
        float outsideDist = NumFast::sqrt_(fmaf(o0, o0, fmaf(o1, o1, o2 * o2)));
        float insideDist = fminf(fmaxf(q0, fmaxf(q1, q2)), 0.f);
        return outsideDist + insideDist;
    } else {
        float qx = NumFast::sqrt_(fmaf(lx, lx, lz * lz)) - pr.x;
        return NumFast::sqrt_(fmaf(qx, qx, ly * ly)) - pr.y;
    }
}
// Fast model, translation-only sphere: local = p + t.
// `rec1` = one float4 (tx,ty,tz,r) per primitive (the streamed search uses the chunk-SoA copy instead).
RM_DEV float prim_sdf_fast_tsphere(const float4* __restrict__ rec1, int j, float x, float y, float z) {
    const float4 s = __ldg(rec1 + j);
    float lx = x + s.x, ly = y + s.y, lz = z + s.z;
    return NumFast::sqrt_(fmaf(lx, lx, fmaf(ly, ly, lz * lz))) - s.w;
}

template <class NP>
struct Ray;
// Exact-build-only evaluation counters (the fast kernels derive the torus count and have no operator trees).
template <bool kExact>
struct EvalExtra {};
template <>
struct EvalExtra<true> {
    unsigned nTorus, opFlops;
    bool fastq;       // this query is evaluated by the fp32 object evaluators (default contexts, march steps)
    bool forceExact;  // ... unless its fp32 result landed next to the hit threshold: the same query is repeated exactly
};

// One scene-object evaluation at the f32 sample point q; r.nSphere / r.nBox (/ r.ex) count primitive evaluations by type.
template <class NP, int PK>
RM_DEV typename NP::F prim_sdf(const RenderParams& P, int j, const float q[3], Ray<NP>& r) {
    unsigned& nSphere = r.nSphere;
    unsigned& nBox = r.nBox;
    if constexpr (NP::kExact) {
        if (P.scene.n_instrs > 0) {
            const unsigned h = P.scene.obj_hist[j];
            nSphere += h & 255u;
            nBox += (h >> 8) & 255u;
            r.ex.nTorus += (h >> 16) & 255u;
            r.ex.opFlops += P.scene.obj_flops[j];
        } else {
            int type = P.scene.type[j];
            nSphere += (type == RM_PRIM_SPHERE);
            nBox += (type == RM_PRIM_BOX);
            r.ex.nTorus += (type == RM_PRIM_TORUS);
        }
        if (r.ex.fastq) return (double)prim_sdf_f32_object(P.scene, j, q[0], q[1], q[2]);
        return prim_sdf_exact(P.scene, j, (double)q[0], (double)q[1], (double)q[2], P.length_sqrt);
    } else if constexpr (PK == PK_TSPHERE) {
        nSphere += 1;
        return prim_sdf_fast_tsphere(P.scene.rec1, j, q[0], q[1], q[2]);
    } else {
        int type;
        float d = prim_sdf_fast_general(P.scene.rec, j, q[0], q[1], q[2], type);
        nSphere += (type == RM_PRIM_SPHERE);
        nBox += (type == RM_PRIM_BOX);
        return d;
    }
}

// The fp64 "polish" of the fast path: the reference's value of primitive j's SDF at q, to double rounding
// (plain sqrt instead of V8's Kahan hypot; the fast path only accepts affine transforms, for which the
// homogeneous divide is by exactly 1).  Keeps the reference's f32 rounding of the local position.
template <int PK>
RM_DEV double prim_sdf_polish(const DevScene& sc, int j, const float q[3]) {
    if constexpr (PK == PK_TSPHERE) {
        const float4 s = __ldg(sc.rec1 + j);
        // f32(x + m12): the float add IS the correctly rounded double sum of two floats
        const double lx = (double)__fadd_rn(q[0], s.x), ly = (double)__fadd_rn(q[1], s.y), lz = (double)__fadd_rn(q[2], s.z);
        return sqrt(lx * lx + ly * ly + lz * lz) - sc.params[4 * (size_t)j];
    } else {
        return leaf_sdf_exact(sc, j, (double)q[0], (double)q[1], (double)q[2], 1);
    }
}

// fp32 evaluation of primitive j by either fast record layout.
template <int PK>
RM_DEV float prim_sdf_f32(const float4* __restrict__ rec, int j, const float q[3]) {
    if constexpr (PK == PK_TSPHERE) {
        return prim_sdf_fast_tsphere(rec, j, q[0], q[1], q[2]);
    } else {
        int type;
        return prim_sdf_fast_general(rec, j, q[0], q[1], q[2], type);
    }
}

// ------------------------------------------------------------------------------------------
// TMA bulk staging of primitive records into shared memory (per-warp double buffer)
// ------------------------------------------------------------------------------------------
constexpr int kStageBytes = 2560;  // one stage = 4 sphere chunks of 640 B (tx,ty,tz,r,|t|^2 x 32) or 40 general records
constexpr int kChunkBytes = 640, kChunkF4 = kChunkBytes / 16;
constexpr int kWarpsPerCtaMax = 16;  // fast BVH kernels: 16 warps per CTA for translation-only spheres (tensor-core pass, 128-request batches),
                                     // 8 for general primitives (64-request batches, 2 per lane); all other kernels 4
constexpr int kChunk = 32;         // argmin granularity of the fp32 search

RM_DEV unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
RM_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
RM_DEV void mbar_init(uint32_t bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
RM_DEV void mbar_expect_tx(uint32_t bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
RM_DEV void bulk_g2s(uint32_t dst, const void* src, unsigned bytes, uint32_t bar) {  // SASS: UBLKCP
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
RM_DEV void mbar_wait(uint32_t bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// Per-warp staging state: two 2 KB shared-memory stages + two mbarriers, phase parity kept in a register.
struct WarpStage {
    const float4* buf[2];
    uint32_t bufAddr[2], bar[2];
    unsigned phase;
    bool resident;  // the whole scene fits one stage and was loaded once at kernel start
};

RM_DEV float4 lds128(uint32_t addr) {  // explicit shared-space load: the compiler cannot fall back to generic LD
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// Blackwell packed fp32 (FADD2 / FMUL2 / FFMA2): two lanes per issue slot.
typedef unsigned long long f32x2;
RM_DEV f32x2 pk2(float a, float b) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
RM_DEV void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
RM_DEV f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
RM_DEV f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
RM_DEV f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// Screening value of the 32 spheres of one chunk-SoA block in shared memory for NQ query points per lane:
// |p + t|^2 - |p|^2 = 2 p.t + |t|^2 with |t|^2 precomputed per sphere — three FFMA2 per pair of spheres per point,
// no MUFU, radii not loaded.  Register tiling over NQ points amortises every broadcast LDS.128 (shared-memory
// wavefronts, not FMAs, are the scarce resource of a one-point loop).  The cancellation error (a few ulp of
// |p|^2 + |t|^2) is covered by the slack of the screen; winners are re-evaluated with the plain formula.
template <int NQ>
RM_DEV void chunk_min_dot(uint32_t ch, const float (&a2)[NQ][3], float (&m)[NQ]) {
#pragma unroll
    for (int k = 0; k < NQ; ++k) m[k] = 3.0e38f;
#pragma unroll 2
    for (int g = 0; g < 8; ++g) {
        const float4 X = lds128(ch + 16u * g), Y = lds128(ch + 128u + 16u * g), Z = lds128(ch + 256u + 16u * g),
                     T = lds128(ch + 512u + 16u * g);
        const f32x2 x0 = pk2(X.x, X.y), x1 = pk2(X.z, X.w), y0 = pk2(Y.x, Y.y), y1 = pk2(Y.z, Y.w);
        const f32x2 z0 = pk2(Z.x, Z.y), z1 = pk2(Z.z, Z.w), t0 = pk2(T.x, T.y), t1 = pk2(T.z, T.w);
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
            const f32x2 ax = pk2(a2[k][0], a2[k][0]), ay = pk2(a2[k][1], a2[k][1]), az = pk2(a2[k][2], a2[k][2]);
            f32x2 s0 = fma2(ax, x0, fma2(ay, y0, fma2(az, z0, t0)));
            f32x2 s1 = fma2(ax, x1, fma2(ay, y1, fma2(az, z1, t1)));
            float a, b, c, d;
            upk2(s0, a, b);
            upk2(s1, c, d);
            m[k] = fminf(m[k], fminf(a, b));
            m[k] = fminf(m[k], fminf(c, d));
        }
    }
}
// Exact fp32 SDFs of the 32 spheres of a chunk held in shared memory: running (min, argmin).
RM_DEV void chunk_exact_smem(uint32_t ch, int base, int valid, const float q[3], float& best, int& idx) {
    const int groups = (valid + 3) >> 2;  // padded dummies inside the last group lose every comparison
#pragma unroll 2
    for (int g = 0; g < groups; ++g) {
        const float4 X = lds128(ch + 16u * g), Y = lds128(ch + 128u + 16u * g), Z = lds128(ch + 256u + 16u * g),
                     R = lds128(ch + 384u + 16u * g);
        const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w}, rs[4] = {R.x, R.y, R.z, R.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float lx = q[0] + xs[u], ly = q[1] + ys[u], lz = q[2] + zs[u];
            float d = NumFast::sqrt_(fmaf(lx, lx, fmaf(ly, ly, lz * lz))) - rs[u];
            if (d < best) {
                best = d;
                idx = base + 4 * g + u;
            }
        }
    }
}
// Same from global memory (per-lane chunk; used to resolve the few screened candidates).
RM_DEV void chunk_exact_gmem(const float4* __restrict__ rec, int chunk, const float q[3], float& best, int& idx) {
    const float4* c = rec + (size_t)chunk * kChunkF4;
#pragma unroll 2
    for (int g = 0; g < 8; ++g) {
        const float4 X = __ldg(c + g), Y = __ldg(c + 8 + g), Z = __ldg(c + 16 + g), R = __ldg(c + 24 + g);
        const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w}, rs[4] = {R.x, R.y, R.z, R.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float lx = q[0] + xs[u], ly = q[1] + ys[u], lz = q[2] + zs[u];
            float d = NumFast::sqrt_(fmaf(lx, lx, fmaf(ly, ly, lz * lz))) - rs[u];
            if (d < best) {
                best = d;
                idx = chunk * 32 + 4 * g + u;
            }
        }
    }
}

// All-primitives search for translation-only spheres over the stages c = first, first + stride, ..., for NQ query
// points per lane (NQ = 2 in the CTA-cooperative pass: 64 requests per batch, lane i takes requests i and i + 32).
//   n_chunks <  kScreenMinChunks : every sphere's SDF (sqrt) straight from the stage;
//   otherwise SCREENED            : the hot loop only computes squared centre distances (chunk_min_dot).  With
//     S = min_j |p - c_j|^2 the nearest surface satisfies d* <= sqrt(S) - r_min, so a sphere of chunk C can only win
//     if |p - c_j| <= sqrt(S) - r_min + r_max(C); chunks passing that test against the RUNNING S are remembered
//     (a handful per query) and resolved exactly afterwards.  Same result as evaluating every SDF.
constexpr int kScreenMinChunks = 16;
constexpr int kCandCap = 40;
template <int NQ>
RM_DEV void search_stages_ts(const RenderParams& P, const float (&q)[NQ][3], WarpStage& ws, int lane, int first, int stride,
                             float (&best)[NQ], int (&code)[NQ]) {
    const float4* __restrict__ rec = P.scene.rec;
    const int nChunks = P.scene.n_chunks;
    const int nStages = (nChunks + 3) / 4;  // 4 chunks (2.5 KB) per stage
    const int myStages = nStages > first ? (nStages - first + stride - 1) / stride : 0;
    auto issue = [&](int i) {
        const int c = first + i * stride, sgi = i & 1;
        const int cnt = min(4, nChunks - 4 * c);
        const unsigned bytes = (unsigned)(cnt * kChunkBytes);
        mbar_expect_tx(ws.bar[sgi], bytes);
        bulk_g2s(ws.bufAddr[sgi], rec + (size_t)c * 4 * kChunkF4, bytes, ws.bar[sgi]);
    };
    const bool resident = ws.resident && first == 0 && stride == 1;
    if (resident) {
        // tiny scenes (<= 4 chunks, staged once per kernel): every sphere's SDF straight from shared memory, no pipeline,
        // no screen — the per-query set-up of the streamed search would cost more than the handful of evaluations
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
            best[k] = 10.f;
            code[k] = -1;
        }
        for (int c = 0; c < nChunks; ++c) {
            const uint32_t ch = ws.bufAddr[0] + (unsigned)kChunkBytes * (unsigned)c;
#pragma unroll
            for (int k = 0; k < NQ; ++k) chunk_exact_smem(ch, c * 32, min(32, P.scene.n_prims - c * 32), q[k], best[k], code[k]);
        }
        return;
    }
    float a2[NQ][3], qq[NQ], E[NQ];
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
        a2[k][0] = 2.f * q[k][0];
        a2[k][1] = 2.f * q[k][1];
        a2[k][2] = 2.f * q[k][2];
        qq[k] = fmaf(q[k][0], q[k][0], fmaf(q[k][1], q[k][1], q[k][2] * q[k][2]));
        // absolute error bound of the screening value: ~4 ulp of the largest intermediate (|p| + |t|max)^2
        E[k] = 5.0e-7f * (qq[k] + P.scene.tt_max + 2.f * NumFast::sqrt_(qq[k] * P.scene.tt_max)) + 1e-30f;
    }
    const float rMin = P.scene.r_min;
    // attempt 0: screened (large scenes); attempt 1: plain SDF of every sphere, taken by the whole warp when some
    // lane's candidate list overflowed (radius spread inside the chunks too wide for the screen to be selective)
    for (int attempt = (nChunks >= kScreenMinChunks) ? 0 : 1; attempt < 2; ++attempt) {
        const bool screened = attempt == 0;
        if (lane == 0 && !resident) {
            if (myStages > 0) issue(0);
            if (myStages > 1) issue(1);
        }
        float sRun[NQ], rootS[NQ];  // running min squared centre distance and (an upper bound of) its square root
        unsigned short candChunk[NQ][kCandCap];
        float candS[NQ][kCandCap];
        int nCand[NQ];
        bool overflow = false;
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
            best[k] = 10.f;
            code[k] = -1;
            sRun[k] = 3.0e38f;
            rootS[k] = 1.0e19f;
            nCand[k] = 0;
        }
        for (int i = 0; i < myStages; ++i) {
            const int sgi = i & 1;
            if (!resident) {
                mbar_wait(ws.bar[sgi], (ws.phase >> sgi) & 1u);
                ws.phase ^= (1u << sgi);
            }
            const uint32_t st = ws.bufAddr[sgi];
            const int chunk0 = 4 * (first + i * stride);
            const int cnt = min(4, nChunks - chunk0);
            // largest radius of each of the stage's 4 chunks (array padded to a multiple of 4 on the host)
            const float4 rm4 = __ldg(reinterpret_cast<const float4*>(P.scene.chunk_rmax) + (first + i * stride));
            const float rmv[4] = {rm4.x, rm4.y, rm4.z, rm4.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c >= cnt) break;
                const uint32_t ch = st + (unsigned)kChunkBytes * (unsigned)c;
                if (!screened) {
#pragma unroll
                    for (int k = 0; k < NQ; ++k)
                        chunk_exact_smem(ch, (chunk0 + c) * 32, min(32, P.scene.n_prims - (chunk0 + c) * 32), q[k], best[k], code[k]);
                } else {
                    float mq[NQ];
                    chunk_min_dot<NQ>(ch, a2, mq);
#pragma unroll
                    for (int k = 0; k < NQ; ++k) {
                        const float m = mq[k] + qq[k];  // min squared centre distance of the chunk, +-E
                        if (m < sRun[k]) {              // new nearest centre
                            sRun[k] = m;
                            rootS[k] = NumFast::sqrt_(fmaxf(m + E[k], 0.f));
                        }
                        const float t = rootS[k] + (rmv[c] - rMin);
                        if (m - E[k] <= t * t * 1.00001f) {
                            if (nCand[k] == kCandCap) {  // compact against the current bound before giving up
                                int w = 0;
                                for (int e = 0; e < nCand[k]; ++e) {
                                    const float te = rootS[k] + (__ldg(P.scene.chunk_rmax + candChunk[k][e]) - rMin);
                                    if (candS[k][e] - E[k] <= te * te * 1.00001f) {
                                        candChunk[k][w] = candChunk[k][e];
                                        candS[k][w] = candS[k][e];
                                        ++w;
                                    }
                                }
                                nCand[k] = w;
                            }
                            if (nCand[k] < kCandCap) {
                                candChunk[k][nCand[k]] = (unsigned short)(chunk0 + c);
                                candS[k][nCand[k]] = m;
                                ++nCand[k];
                            } else {
                                overflow = true;
                            }
                        }
                    }
                }
            }
            __syncwarp();  // every lane is done reading this stage before it is refilled
            if (lane == 0 && i + 2 < myStages) issue(i + 2);
        }
        if (!screened) break;
        if (__any_sync(kFull, overflow)) continue;  // rare: redo the pass unscreened, all lanes together
#pragma unroll
        for (int k = 0; k < NQ; ++k)
            for (int e = 0; e < nCand[k]; ++e) {
                const float te = rootS[k] + (__ldg(P.scene.chunk_rmax + candChunk[k][e]) - rMin);
                if (candS[k][e] - E[k] <= te * te * 1.00001f) chunk_exact_gmem(rec, (int)candChunk[k][e], q[k], best[k], code[k]);
            }
        break;
    }
    // padded dummies can never win (|l| ~ 1e15), so code < n_prims whenever it is set
}

// ------------------------------------------------------------------------------------------
// Tensor-core search (tcgen05 + TMEM) for translation-only spheres in the CTA-cooperative pass.
//
// The screening value of the search is a distance matrix: S[i][j] = |p_i + t_j|^2 - |p_i|^2 = 2 p_i . t_j + |t_j|^2 for
// 128 query points x all spheres — a GEMM with K = 4.  It runs on the 5th-generation tensor cores in TF32 with the
// fp32 operands SPLIT into hi + lo TF32 terms (hi.hi + lo.hi + hi.lo + lo.lo per coordinate, |t|^2 in three terms:
// K = 16), so the products are exact in the fp32 accumulator and the result is good to ~1e-7 of (|p| + |t|)^2 — the same
// error class as the FFMA version it replaces, covered by the screen's slack.  Operands are K-major, no swizzle:
// the B tiles (128 spheres x 16) are pre-tiled in HBM as exact shared-memory images and arrive by TMA bulk copy;
// accumulators live in TMEM (2 x 128 columns, double buffered); the epilogue pulls them with tcgen05.ld and keeps
// one minimum per 32-sphere chunk (FMNMX3), feeding the same candidate screen as before.  Results are unchanged:
// candidates are still resolved with the plain fp32 SDF and the winner polished in fp64.
// ------------------------------------------------------------------------------------------
constexpr int kTcGroups = 4;            // warpgroups of the CTA = TMEM accumulator buffers (4 x 128 columns = all of TMEM)
#ifndef RM_TC_STAGES
#define RM_TC_STAGES 8
#endif
constexpr int kTcStages = RM_TC_STAGES;  // B-tile ring (4 KB each)
#ifndef RM_TC_ITEM_CAP
#define RM_TC_ITEM_CAP 7168  // 28 KB: with the static cooperative queue the tensor-core instance then needs 68 KB of dynamic + 25 KB of
                             // static shared memory — under the 100 KB carve-out step, which leaves the SM 156 KB of L1 instead of 124
#endif
constexpr unsigned kTcItemCap = RM_TC_ITEM_CAP;  // work items of one pass (a typical pass lists ~550; overflow falls back to a per-lane scan)
constexpr int kTcBlock = 128;           // spheres per MMA (N) = query rows per batch (M)
constexpr uint32_t kTcTileBytes = 4096;  // B tile: 128 rows x 8 tf32
constexpr uint32_t kTcATileBytes = 8192;  // A tile: 128 rows x 16 tf32
constexpr uint32_t kTcLBO = 2048, kTcSBO = 128;  // canonical no-swizzle K-major: 8-row core matrices of 128 B

RM_DEV uint64_t umma_desc(uint32_t saddr) {  // shared-memory matrix descriptor (sm_100, version 1, SWIZZLE_NONE)
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(kTcLBO >> 4) << 16) | ((uint64_t)(kTcSBO >> 4) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor: kind::tf32, fp32 accumulate, A and B K-major, M = N = 128
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcBlock >> 3) << 17) | ((uint32_t)(kTcBlock >> 4) << 24);
RM_DEV void umma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t accumulate) {  // SASS: UTCHMMA
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a), "l"(b), "r"(kTcIdesc), "r"(accumulate)
        : "memory");
}
RM_DEV void umma_commit(uint32_t bar) {  // SASS: UTCBAR — the mbarrier completes when every MMA issued so far has
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
RM_DEV void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
RM_DEV void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {  // SASS: LDTM.x32 — 32 columns of this thread's TMEM lane
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
        "%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr));
}
RM_DEV void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {  // SASS: LDTM.x16
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
// min3 as a volatile asm: keeps its place between the (volatile) TMEM loads in program order, so the reduction of one
// register set sits between the issue of the next loads — the hardware scoreboard releases each set as it lands.
RM_DEV float min3v(float a, uint32_t b, uint32_t c) {
    float r;
    asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "r"(b), "r"(c));
    return r;
}
// minimum of 32 accumulators held in two 16-register sets; every instruction mixes both sets
RM_DEV float min16x2(const uint32_t (&a)[16], const uint32_t (&b)[16]) {
    float m0 = min3v(__uint_as_float(a[0]), b[0], a[1]);
    float m1 = min3v(__uint_as_float(b[1]), a[2], b[2]);
    float m2 = min3v(__uint_as_float(a[3]), b[3], a[4]);
    float m3 = min3v(__uint_as_float(b[4]), a[5], b[5]);
    m0 = min3v(m0, a[6], b[6]);
    m1 = min3v(m1, a[7], b[7]);
    m2 = min3v(m2, a[8], b[8]);
    m3 = min3v(m3, a[9], b[9]);
    m0 = min3v(m0, a[10], b[10]);
    m1 = min3v(m1, a[11], b[11]);
    m2 = min3v(m2, a[12], b[12]);
    m3 = min3v(m3, a[13], b[13]);
    m0 = min3v(m0, a[14], b[14]);
    m1 = min3v(m1, a[15], b[15]);
    return fminf(fminf(m0, m1), fminf(m2, m3));
}
RM_DEV float min3(float a, float b, float c) {  // SASS: FMNMX3
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
RM_DEV float min32(const uint32_t (&v)[32]) {
    float a = min3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
    float b = min3(__uint_as_float(v[3]), __uint_as_float(v[4]), __uint_as_float(v[5]));
    float c = min3(__uint_as_float(v[6]), __uint_as_float(v[7]), __uint_as_float(v[8]));
    float d = min3(__uint_as_float(v[9]), __uint_as_float(v[10]), __uint_as_float(v[11]));
#pragma unroll
    for (int j = 12; j < 28; j += 8) {
        a = min3(a, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
        b = min3(b, __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        c = min3(c, __uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
        d = min3(d, __uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
    }
    a = min3(a, __uint_as_float(v[28]), __uint_as_float(v[29]));
    b = min3(b, __uint_as_float(v[30]), __uint_as_float(v[31]));
    return fminf(min3(a, b, c), d);
}
RM_DEV float tf32_rna(float x) {  // round to TF32 (nearest, ties away): the low 13 mantissa bits end up zero
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// Per-CTA tensor-core state (shared-memory addresses + the running block counter that fixes every mbarrier parity).
struct TcCtx {
    uint32_t tmem;       // TMEM base: 512 columns = four 128-column accumulator buffers (one per warpgroup)
    uint32_t sA, sB;     // shared-memory addresses: A tile (8 KB), B ring (kTcStages x 4 KB)
    uint32_t barFull;    // kTcStages mbarriers: B tile landed
    uint32_t barTmemFull;  // 4 mbarriers: accumulator buffer written
    uint32_t drainCnt;     // 4 counters: warps of the owning group that have drained the buffer
    unsigned g;          // blocks processed by this CTA so far (uniform across the CTA)
};

// One cooperative pass: the exact nearest-surface distance over EVERY sphere for up to 128 requests.  Called convergently
// by all 512 threads of the CTA.
//
// 1. Cluster screen on the tensor cores (see rm_api.cu): S_j = |p - C_j|^2 for 128 queries x all clusters of 128 spheres.
//    Thread `row` = threadIdx.x & 127 owns TMEM lane / query `row`; the four 4-warp groups `grp` = threadIdx.x >> 7 split the
//    128-cluster blocks ((g0 + b) % 4 == grp, accumulator buffer `grp`).  From its accumulators a thread keeps
//      ub = min_j (sqrt(S_j) + u_j)           — some sphere is at most this far,
//      the clusters with sqrt(S_j) - R_j <= ub — the only ones that can hold the nearest sphere.
// 2. ub is combined over the four groups; the surviving (query, cluster) pairs become work items in shared memory.
// 3. The 16 warps drain the items: a warp loads the cluster's 128 spheres (sorted chunk-SoA copy, coalesced), evaluates the plain
//    fp32 SDF at the item's query, and folds (distance, index) into the query's 64-bit key with one shared-memory atomicMin.
// The result (distance, scene primitive index) of query `row` goes to partBest / partCode [row].
//
// Operand packing (K = 16 in two MMAs that read the SAME 8-wide B tile):
//   B row (8 tf32)   : t_hi.x t_hi.y t_hi.z |t|^2_hi   t_lo.x t_lo.y t_lo.z |t|^2_lo          (t = -C_j)
//   A row, MMA 1     : 2p_hi.x 2p_hi.y 2p_hi.z 1        2p_hi.x 2p_hi.y 2p_hi.z 1      -> 2 p_hi.(t_hi + t_lo) + |t|^2
//   A row, MMA 2     : 2p_lo.x 2p_lo.y 2p_lo.z 0        2p_lo.x 2p_lo.y 2p_lo.z 0      -> 2 p_lo.(t_hi + t_lo)
RM_DEV uint32_t float_flip(float f) {  // order-preserving map float -> uint32
    const uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
RM_DEV float float_unflip(uint32_t k) {
    return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}
static __device__ __noinline__ void tc_pass(const RenderParams& P, TcCtx& tcRef, const float4* shReq, unsigned head, unsigned nBatch, int qcap,
                                            float* partBest, int* partCode, unsigned long long* shKey, unsigned* shItemCount) {
    const TcCtx tc = tcRef;  // registers: the asm memory clobbers below would otherwise force reloads from local memory
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wq = warp & 3;
    const int row = tid & 127, grp = tid >> 7;
    const bool sweeper = tid < 4 * kTcBlock;  // the first 16 warps: four 4-warp groups; any further warps only drain work items
    const bool valid = sweeper && (unsigned)row < nBatch;
#ifdef RM_PHASE_TIMING
    long long tcT[4] = {0, 0, 0, 0}, tcMark = clock64();
#define RM_TC_MARK(i)                     \
    do {                                  \
        const long long now_ = clock64(); \
        tcT[i] += now_ - tcMark;          \
        tcMark = now_;                    \
    } while (0)
#else
#define RM_TC_MARK(i)
#endif
    float q[3] = {0.f, 0.f, 0.f};
    if (valid) {
        const float4 v = shReq[(head + (unsigned)row) % (unsigned)qcap];
        q[0] = v.x;
        q[1] = v.y;
        q[2] = v.z;
    }
    if (grp == 0) {  // A tile, canonical no-swizzle K-major: (row, K chunk c) at c * LBO + (row / 8) * SBO + (row % 8) * 16
        float hi[3], lo[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float p2 = 2.f * q[c];
            hi[c] = tf32_rna(p2);
            lo[c] = tf32_rna(p2 - hi[c]);
        }
        const float one = valid ? 1.f : 0.f;
        const uint32_t base = tc.sA + (uint32_t)(row >> 3) * kTcSBO + (uint32_t)(row & 7) * 16u;
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base), "f"(hi[0]), "f"(hi[1]), "f"(hi[2]), "f"(one) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + kTcLBO), "f"(hi[0]), "f"(hi[1]), "f"(hi[2]), "f"(one) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + 2u * kTcLBO), "f"(lo[0]), "f"(lo[1]), "f"(lo[2]), "f"(0.f) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + 3u * kTcLBO), "f"(lo[0]), "f"(lo[1]), "f"(lo[2]), "f"(0.f) : "memory");
        shKey[row] = ((unsigned long long)float_flip(3.0e38f) << 32) | 0xFFFFFFFFull;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core (async proxy)
    __syncthreads();

    const int nB = P.scene.n_tc_blocks, nCl = P.scene.n_clusters;
    const char* tiles = reinterpret_cast<const char*>(P.scene.tc_tiles);
    const float2* __restrict__ bound = P.scene.cl_bound;
    unsigned gBase = tc.g;  // global block counter at the start of the current sweep (fixes all mbarrier parities)
    // Scenes of <= kTcStages cluster blocks (128 x 128 x 8 = 131 072 spheres): every B tile has its own stage, loaded once at
    // kernel start and resident for the whole frame — a pass then starts its sweeps without a TMA round trip (~2.5 us under load,
    // twice per pass).  Bigger scenes stream the tiles through the ring as before.
    const bool resB = RM_RESIDENT_B && nB <= kTcStages;
    auto tma = [&](int b) {  // B tile b -> its ring stage
        if (resB) return;
        const unsigned sidx = (gBase + (unsigned)b) % kTcStages;
        mbar_expect_tx(tc.barFull + 8u * sidx, kTcTileBytes);
        bulk_g2s(tc.sB + sidx * kTcTileBytes, tiles + (size_t)b * kTcTileBytes, kTcTileBytes, tc.barFull + 8u * sidx);
    };
    auto mma = [&](int b) {  // S[128 x 128] of block b -> the accumulator buffer of its group
        const unsigned g = gBase + (unsigned)b, sidx = resB ? (unsigned)b : g % kTcStages;
        // (the accumulator buffer is free: the caller is the last of the group's warps to have drained block g - 4)
        if (!resB) mbar_wait(tc.barFull + 8u * sidx, (g / kTcStages) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tc.tmem + (g % kTcGroups) * (uint32_t)kTcBlock;
        const uint64_t db = umma_desc(tc.sB + sidx * kTcTileBytes);
        umma_tf32(d, umma_desc(tc.sA), db, 0u);
        umma_tf32(d, umma_desc(tc.sA + 2u * kTcLBO), db, 1u);
        umma_commit(tc.barTmemFull + 8u * (g % kTcGroups));
    };

    const float qq = fmaf(q[0], q[0], fmaf(q[1], q[1], q[2] * q[2]));
    // error bound of the split-TF32 value: measured ~1e-7 of (|p| + |t|)^2 (tools/tc_test.cu); 2e-6 budgeted
    const float E = 2.0e-6f * (qq + P.scene.tt_max + 2.f * NumFast::sqrt_(qq * P.scene.tt_max)) + 1e-30f;
    const uint32_t taddr = tc.tmem + (uint32_t)grp * (uint32_t)kTcBlock + ((uint32_t)(wq * 32) << 16);
    const uint32_t barTFullG = tc.barTmemFull + 8u * (unsigned)grp, drainG = tc.drainCnt + 4u * (unsigned)grp;
    uint32_t* items;  // work-item list (row << 16 | cluster): the dynamic shared memory behind the A tile and the B ring
    {
        extern __shared__ __align__(128) float4 shDynTc[];
        items = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(shDynTc) + kTcATileBytes + kTcStages * kTcTileBytes);
    }
    constexpr unsigned kItemCap = kTcItemCap;

    // One sweep of the tensor cores over all cluster blocks; `consume(b, h, v0, v1)` gets 64 accumulators of block b at a time.
    auto sweep = [&](auto&& consume) {
        if (tid == 0) {
            for (int b = 0; b < kTcStages && b < nB; ++b) tma(b);
            for (int b = 0; b < kTcGroups && b < nB; ++b) mma(b);
        }
        __syncwarp();
        const int b0 = sweeper ? (int)(((unsigned)grp - gBase) % (unsigned)kTcGroups) : nB;  // this group's first block: (gBase + b0) % 4 == grp
        for (int b = b0; b < nB; b += kTcGroups) {
            const unsigned u = (gBase + (unsigned)b) / kTcGroups;
            mbar_wait(barTFullG, u & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                uint32_t v0[32], v1[32];
                tmem_ld32(taddr + 64u * (unsigned)h, v0);
                tmem_ld32(taddr + 64u * (unsigned)h + 32u, v1);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (h == 1) {
                    // the accumulator buffer is free as soon as its values are in registers; whichever of the group's four
                    // warps drains it LAST re-arms it: no warp ever waits for a straggler
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        unsigned old;
                        asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(drainG) : "memory");
                        if (old == 3u) {
                            asm volatile("st.relaxed.cta.shared::cta.u32 [%0], %1;" ::"r"(drainG), "r"(0u) : "memory");
                            if (b + kTcGroups < nB) mma(b + kTcGroups);    // next block of this buffer
                            if (b + kTcStages < nB) tma(b + kTcStages);  // MMA b has completed: its B stage is free
                        }
                    }
                    __syncwarp();
                }
                consume(b, h, v0, v1);
            }
        }
        __syncthreads();  // every MMA of the sweep has completed and been drained; the B ring is idle
        gBase += (unsigned)nB;
    };
    // Work items: one warp per (query, cluster) pair, 4 spheres per lane, coalesced loads from the sorted chunk-SoA copy;
    // (distance, index) folded into the query's key with one shared-memory atomicMin.  Two items in flight per warp.
    auto drain_items = [&](unsigned nItems) {
        const float* __restrict__ cr = reinterpret_cast<const float*>(P.scene.cl_rec);
        auto eval = [&](uint32_t item, float& bestD, unsigned& bestI, unsigned& irow) {
            irow = item >> 16;
            const unsigned cl = item & 0xFFFFu;
            const float4 qv = shReq[(head + irow) % (unsigned)qcap];
            bestD = 3.0e38f;
            bestI = 0xFFFFFFFFu;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float* ch = cr + (size_t)(cl * 4u + (unsigned)c) * 160u + (unsigned)lane;
                const float lx = qv.x + __ldg(ch), ly = qv.y + __ldg(ch + 32), lz = qv.z + __ldg(ch + 64);
                const float d = NumFast::sqrt_(fmaf(lx, lx, fmaf(ly, ly, lz * lz))) - __ldg(ch + 96);
                if (d < bestD) {
                    bestD = d;
                    bestI = (cl * 4u + (unsigned)c) * 32u + (unsigned)lane;
                }
            }
        };
        auto fold = [&](float bestD, unsigned bestI, unsigned irow) {
            const uint32_t key = float_flip(bestD);
            const uint32_t kmin = __reduce_min_sync(kFull, key);
            const uint32_t imin = __reduce_min_sync(kFull, key == kmin ? bestI : 0xFFFFFFFFu);
            if (lane == 0) atomicMin(&shKey[irow], ((unsigned long long)kmin << 32) | (unsigned long long)imin);
        };
        const unsigned nWarps = blockDim.x >> 5;
        for (unsigned it = (unsigned)warp; it < nItems; it += 2u * nWarps) {
            float d0, d1 = 3.0e38f;
            unsigned i0, i1 = 0xFFFFFFFFu, r0, r1 = 0u;
            const bool two = it + nWarps < nItems;
            eval(items[it], d0, i0, r0);
            if (two) eval(items[it + nWarps], d1, i1, r1);
            fold(d0, i0, r0);
            if (two) fold(d1, i1, r1);
        }
        __syncthreads();
    };

    // ---- 1. first sweep: the nearest cluster centre of every query (pure minimum: FMNMX3), then that cluster's exact SDFs
    float sMin = 3.0e38f;
    int jMin = 0;
    RM_TC_MARK(3);
    sweep([&](int b, int h, const uint32_t(&v0)[32], const uint32_t(&v1)[32]) {
        const float m = fminf(min32(v0), min32(v1));
        if (m < sMin) {  // rare after the first blocks: locate the column
            sMin = m;
            int k = 0;
#pragma unroll
            for (int i = 31; i >= 0; --i) {
                if (__uint_as_float(v1[i]) == m) k = 32 + i;
                if (__uint_as_float(v0[i]) == m) k = i;
            }
            jMin = b * kTcBlock + 64 * h + k;
        }
    });
    RM_TC_MARK(0);
    if (sweeper) {
        partBest[grp * 128 + row] = sMin;
        partCode[grp * 128 + row] = jMin;
    }
    __syncthreads();
    if (grp == 0) {
#pragma unroll
        for (int g2 = 1; g2 < kTcGroups; ++g2)
            if (partBest[g2 * 128 + row] < sMin) {
                sMin = partBest[g2 * 128 + row];
                jMin = partCode[g2 * 128 + row];
            }
        if (valid) items[row] = ((unsigned)row << 16) | (unsigned)min(jMin, nCl - 1);
    }
    if (tid == 0) *shItemCount = 0u;
    __syncthreads();
    drain_items(nBatch);
    // ub: the fp32 SDF of an actual sphere — nothing farther than this can be the nearest
    float ub = float_unflip((uint32_t)(shKey[row] >> 32));  // (row = tid & 127: defined for every thread)
    ub = ub + fabsf(ub) * 1.0e-6f + 1.0e-7f;
    __syncthreads();  // every thread has read the keys / the first item round before the list is rebuilt
    RM_TC_MARK(1);

    // ---- 2. second sweep: clusters with sqrt(S_j) - R_j <= ub, i.e. S_j <= (ub + R_j)^2, become work items
    const float qqLo = qq - E;
    sweep([&](int b, int h, const uint32_t(&v0)[32], const uint32_t(&v1)[32]) {
        if (!valid) return;
        // one threshold for the whole block (its largest R), on the raw accumulator; the rare passers get their own R_j
        const float tb = fmaxf(ub + __ldg(P.scene.cl_block_rmax + b), 0.f);
        const float thr = tb * tb * 1.00001f - qqLo;
        bool any = false;
#pragma unroll
        for (int i = 0; i < 32; ++i) any |= (__uint_as_float(v0[i]) <= thr) | (__uint_as_float(v1[i]) <= thr);
        if (any) {
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                const float sv = __uint_as_float(i < 32 ? v0[i & 31] : v1[i & 31]);
                if (sv <= thr) {
                    const int j = b * kTcBlock + 64 * h + i;
                    const float t = fmaxf(ub + __ldg(bound + j).x, 0.f);
                    if (j < nCl && sv + qqLo <= t * t * 1.00001f) {
                        const unsigned slot = atomicAdd(shItemCount, 1u);
                        if (slot < kItemCap) items[slot] = ((unsigned)row << 16) | (unsigned)j;
                    }
                }
            }
        }
    });
    RM_TC_MARK(0);
    const unsigned nItems = min(*(volatile unsigned*)shItemCount, kItemCap);
    const bool listOverflow = *(volatile unsigned*)shItemCount > kItemCap;
    drain_items(nItems);
    RM_TC_MARK(2);
    tcRef.g = gBase;
    if (tid == 0) {
        atomicAdd(&P.stats->tc_passes, 1ull);
        atomicAdd(&P.stats->tc_requests, (unsigned long long)nBatch);
        atomicAdd(&P.stats->tc_items, (unsigned long long)(nBatch + nItems));
    }
#ifdef RM_PHASE_TIMING
    if (tid == 0) {
        atomicAdd(&P.stats->n_pass, 1ull);
        atomicAdd(&P.stats->n_req, (unsigned long long)nBatch);
        atomicAdd(&P.stats->n_rearm, (unsigned long long)nItems);
        for (int i = 0; i < 4; ++i) atomicAdd(&P.stats->t_tc[i], (unsigned long long)tcT[i]);
    }
#endif
    // results: partial 0 carries the answer, the other three are neutral
    float best = 10.f;
    int code = -1;
    if (grp == 0 && valid) {
        const unsigned long long k = shKey[row];
        uint32_t idx = (uint32_t)(k & 0xFFFFFFFFull);
        float d = float_unflip((uint32_t)(k >> 32));
        if (listOverflow) {  // pathological scene (item list full): every sphere, per lane
            int c2 = -1;
            float b2 = 3.0e38f;
            for (int c = 0; c < nCl * 4; ++c) chunk_exact_gmem(P.scene.cl_rec, c, q, b2, c2);
            if (c2 >= 0) {
                d = b2;
                idx = (uint32_t)c2;
            }
        }
        if (idx != 0xFFFFFFFFu && d < 10.f) {
            best = d;
            code = __ldg(P.scene.cl_perm + idx);
        }
    }
    if (sweeper) {
        partBest[grp * 128 + row] = best;
        partCode[grp * 128 + row] = code;
    }
}

template <int PK>
RM_DEV float sdf_from_stage(uint32_t st, int k, const float q[3]) {
    if constexpr (PK == PK_TSPHERE) {
        const float4 s = lds128(st + 16u * (unsigned)k);
        float lx = q[0] + s.x, ly = q[1] + s.y, lz = q[2] + s.z;
        return NumFast::sqrt_(fmaf(lx, lx, fmaf(ly, ly, lz * lz))) - s.w;
    } else {
        const float4 r0 = lds128(st + 64u * (unsigned)k), r1 = lds128(st + 64u * (unsigned)k + 16u), r2 = lds128(st + 64u * (unsigned)k + 32u),
                     pr = lds128(st + 64u * (unsigned)k + 48u);
        float lx = fmaf(r0.x, q[0], fmaf(r0.y, q[1], fmaf(r0.z, q[2], r0.w)));
        float ly = fmaf(r1.x, q[0], fmaf(r1.y, q[1], fmaf(r1.z, q[2], r1.w)));
        float lz = fmaf(r2.x, q[0], fmaf(r2.y, q[1], fmaf(r2.z, q[2], r2.w)));
        int type = __float_as_int(pr.w);
        if (type == RM_PRIM_SPHERE) {
            return NumFast::sqrt_(fmaf(lx, lx, fmaf(ly, ly, lz * lz))) - pr.x;
        } else if (type == RM_PRIM_BOX) {
            float q0 = fabsf(lx) - pr.x, q1 = fabsf(ly) - pr.y, q2 = fabsf(lz) - pr.z;
            float o0 = fmaxf(q0, 0.f), o1 = fmaxf(q1, 0.f), o2 = fmaxf(q2, 0.f);
            return NumFast::sqrt_(fmaf(o0, o0, fmaf(o1, o1, o2 * o2))) + fminf(fmaxf(q0, fmaxf(q1, q2)), 0.f);
        } else {
            float qx = NumFast::sqrt_(fmaf(lx, lx, lz * lz)) - pr.x;
            return NumFast::sqrt_(fmaf(qx, qx, ly * ly)) - pr.y;
        }
    }
}

// Dense all-primitives evaluation: closest = min over every primitive (scene.ts:183-189), from MAX_DIST = 10.
//   exact model: the reference's Math.min chain in doubles (per lane, records read from global memory);
//   fast model : executed by the WHOLE warp (call it convergently; lanes with active == false just help
//                with the staging).  The primitive array streams through the warp's two shared-memory
//                stages with TMA bulk copies (one elected lane issues cp.async.bulk; completion on an
//                mbarrier) while every lane evaluates its own sample point against the stage in flight —
//                one broadcast LDS.128 per primitive.  fp32 SEARCH for the nearest primitive keeps a
//                running min per 32-primitive chunk (FMNMX only); the winning chunk is re-scanned once
//                to recover the index, and that ONE primitive is evaluated in fp64 (the "polish").
// fp32 search over the stages c = first, first + stride, ... (every warp of a CTA takes an interleaved
// share when the pass is CTA-cooperative; first = 0, stride = 1 when a warp serves itself).
// Returns the running min and a code: idx >= 0 exact index | (chunk | kChunkFlag) 32-primitive chunk | -1 none.
constexpr int kChunkFlag = 0x40000000;
template <int PK>
RM_DEV void search_stages(const RenderParams& P, const float q[3], WarpStage& ws, int lane, int first, int stride, float& best,
                          int& code) {
    if constexpr (PK == PK_TSPHERE) {
        const float q1[1][3] = {{q[0], q[1], q[2]}};
        float b1[1];
        int c1[1];
        search_stages_ts<1>(P, q1, ws, lane, first, stride, b1, c1);
        best = b1[0];
        code = c1[0];
        return;
    }
    constexpr int kF4 = (PK == PK_TSPHERE) ? 1 : 4;     // float4 per primitive
    constexpr int kPerStage = kStageBytes / (16 * kF4);  // primitives per stage (128 / 32)
    const int n = P.scene.n_prims;
    const float4* __restrict__ rec = P.scene.rec;
    const int nStages = (n + kPerStage - 1) / kPerStage;
    const int myStages = nStages > first ? (nStages - first + stride - 1) / stride : 0;
    auto issue = [&](int i) {  // elected lane: my i-th stage -> buffer i&1
        const int c = first + i * stride, sgi = i & 1;
        const int cnt = min(kPerStage, n - c * kPerStage);
        const unsigned bytes = (unsigned)(cnt * 16 * kF4);
        mbar_expect_tx(ws.bar[sgi], bytes);
        bulk_g2s(ws.bufAddr[sgi], rec + (size_t)c * kPerStage * kF4, bytes, ws.bar[sgi]);
    };
    const bool resident = ws.resident && first == 0 && stride == 1;
    if (lane == 0 && !resident) {
        if (myStages > 0) issue(0);
        if (myStages > 1) issue(1);
    }
    best = 10.f;
    code = -1;
    for (int i = 0; i < myStages; ++i) {
        const int sgi = i & 1;
        if (!resident) {
            mbar_wait(ws.bar[sgi], (ws.phase >> sgi) & 1u);
            ws.phase ^= (1u << sgi);
        }
        const uint32_t st = ws.bufAddr[sgi];
        const int base = (first + i * stride) * kPerStage;
        const int cnt = min(kPerStage, n - base);
        int k = 0;
        for (; k + kChunk <= cnt; k += kChunk) {
            float m = sdf_from_stage<PK>(st, k, q);
#pragma unroll
            for (int u = 1; u < kChunk; ++u) m = fminf(m, sdf_from_stage<PK>(st, k + u, q));
            if (m < best) {
                best = m;
                code = (base + k) | kChunkFlag;
            }
        }
        for (; k < cnt; ++k) {  // tail (and the whole scene when n < 32)
            float d = sdf_from_stage<PK>(st, k, q);
            if (d < best) {
                best = d;
                code = base + k;
            }
        }
        __syncwarp();  // every lane is done reading this stage before it is refilled
        if (lane == 0 && i + 2 < myStages) issue(i + 2);
    }
}

// Turn the search result into the scene distance: recover the index inside the winning chunk, then ONE
// fp64 evaluation of that primitive (the "polish"), clamped to MAX_DIST like scene.ts:145-146.
template <int PK, bool kExactIndexOnly = false>
RM_DEV double finish_search(const RenderParams& P, const float q[3], int code) {
    if (code < 0) return 10.0;  // nothing closer than MAX_DIST
    int idx = code;
    if (!kExactIndexOnly && (code & kChunkFlag)) {  // (the cluster screen always returns an exact index)
        const int chunk = code & ~kChunkFlag;
        float mm = 3.0e38f;
        for (int k = 0; k < kChunk; ++k) {
            float d = prim_sdf_f32<PK>((PK == PK_TSPHERE) ? P.scene.rec1 : P.scene.rec, chunk + k, q);
            if (d < mm) {
                mm = d;
                idx = chunk + k;
            }
        }
    }
    return jsmin(prim_sdf_polish<PK>(P.scene, idx, q), 10.0);
}

// Dense all-primitives evaluation by one warp for its own lanes: closest = min over every primitive
// (scene.ts:183-189), from MAX_DIST = 10.
//   exact model: the reference's Math.min chain in doubles (per lane, records read from global memory);
//   fast model : executed by the WHOLE warp (call it convergently; lanes with active == false only help
//                with the staging).  The primitive array streams through the warp's two shared-memory
//                stages with TMA bulk copies (one elected lane issues cp.async.bulk; completion on an
//                mbarrier) while every lane evaluates its own sample point against the stage in flight —
//                one broadcast LDS.128 per primitive.  The fp32 SEARCH keeps only a running min per
//                32-primitive chunk (FMNMX); the winner is re-scanned once for its index and that ONE
//                primitive is evaluated in fp64.
template <class NP, int PK>
RM_DEV double scene_all_prims(const RenderParams& P, const float q[3], WarpStage& ws, bool active, int lane, bool fastq = false) {
    if constexpr (NP::kExact) {
        const int n = P.scene.n_prims;
        double closest = 10.0;
        if (active) {
            if (fastq) {  // default contexts, march steps: fp32 object evaluators
                float cf = 10.f;
                for (int j = 0; j < n; ++j) cf = fminf(cf, prim_sdf_f32_object(P.scene, j, q[0], q[1], q[2]));
                closest = (double)cf;
            } else {
                for (int j = 0; j < n; ++j)
                    closest = jsmin(prim_sdf_exact(P.scene, j, (double)q[0], (double)q[1], (double)q[2], P.length_sqrt), closest);
            }
        }
        return closest;
    } else {
        float best;
        int code;
        search_stages<PK>(P, q, ws, lane, 0, 1, best, code);
        if (!active) return 10.0;
        return finish_search<PK>(P, q, code);
    }
}

// Candidate-list evaluation (octree leaf / BVH leaf): running (min, argmin) in the field model.
template <class NP, int PK>
RM_DEV void leaf_prims(const RenderParams& P, const int32_t* __restrict__ lp, int pc, const float q[3], double& distExact,
                       float& distF32, int& argmin, Ray<NP>& r) {
    for (int k = 0; k < pc; ++k) {
        const int j = lp[k];
        if constexpr (NP::kExact) {
            distExact = jsmin(prim_sdf<NP, PK>(P, j, q, r), distExact);
        } else {
            float d = (float)prim_sdf<NP, PK>(P, j, q, r);
            if (d < distF32) {
                distF32 = d;
                argmin = j;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Bounding boxes.  Points and box corners are both float32 VALUES, so containment is exact in float.
// ------------------------------------------------------------------------------------------
RM_DEV bool box_contains(const float* bmin, const float* bmax, const float q[3]) {  // boundingBox.ts:15-21
    return q[0] >= bmin[0] && q[0] <= bmax[0] && q[1] >= bmin[1] && q[1] <= bmax[1] && q[2] >= bmin[2] && q[2] <= bmax[2];
}

// BoundingBox.intersectRay (boundingBox.ts:69-105), doubles.  invD = 1/direction[i] is hoisted out of
// the per-box loop (same value every time).  No NaN can arise: |d| >= 1e-10 keeps invD finite.
RM_DEV bool box_intersect_ray(const float* bmin, const float* bmax, const double o[3], const float d[3], const double invD[3],
                              double& tMinOut, double& tMaxOut) {
    double tMin = -d_inf(), tMax = d_inf();
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (fabs((double)d[i]) < 1e-10) {
            if (o[i] < (double)bmin[i] || o[i] > (double)bmax[i]) return false;
        } else {
            double t0 = ((double)bmin[i] - o[i]) * invD[i];
            double t1 = ((double)bmax[i] - o[i]) * invD[i];
            if (t0 > t1) {
                double tmp = t0;
                t0 = t1;
                t1 = tmp;
            }
            tMin = tMin > t0 ? tMin : t0;  // Math.max, NaN-free
            tMax = tMax < t1 ? tMax : t1;
            if (tMin > tMax) return false;
        }
    }
    tMinOut = tMin;
    tMaxOut = tMax;
    return true;
}

// ------------------------------------------------------------------------------------------
// Per-lane ray state
// ------------------------------------------------------------------------------------------
template <class NP>
struct Ray {
    float d[3];   // unit direction (f32 values, raymarcher.ts:84-88)
    float q[3];   // pending query point (f32 values)
    float h[3];   // hit position (raymarcher.ts:94-95)
    double t;     // totalDist
    double prevSDF, prevStep;  // V2 / V3
    double aux0, aux1;         // V3: originalPos / newSDF
    double depth;              // rayMarch return
    double nd;                 // getNormal: d = scene distance at the hit position
    float n0, n1, n2;          // normal (f32 values)
    unsigned sdf, iters;       // un-wrapped counters
    unsigned nSphere, nBox;    // evaluations by type (fast kernels: torus = sdf - nSphere - nBox)
    EvalExtra<NP::kExact> ex;   // exact kernels: explicit torus count + operator FLOPs
    int i;                     // march loop index
    int phase;
    int px, py;  // pixel (x, band-local y)
    bool done;   // rayMarch has returned (depth is valid)
    bool pending;  // parked on the dense all-primitives pass
    int cur, nIv;  // BVH interval cursor (bvh.ts:204-240)
    double curEnter, curExit;  // the interval under the cursor
    bool curValid, lazy;
};

// BVH per-ray interval list.  The cursor advances at most one slot per march-loop index
// (bvh.ts:222-236) and there are at most MAX_STEPS indices, so only the first MAX_STEPS+1 entries of
// the stably sorted list can ever be read; exactly those are kept, by bounded stable insertion.
struct IvList {
    double enter[kMaxStepsFixed + 1];
    double exit_[kMaxStepsFixed + 1];
};

// BVH.findRayIntersections + onRayMarchStart (bvh.ts:126-202).  Returns the number of kept intervals.
// (Called out of line: the ray origin comes as a pointer into the kernel parameters and the direction BY VALUE, so that no
// address of the per-lane ray state escapes — an escaping pointer would pin the whole Ray struct in local memory.)
static __device__ __noinline__ int bvh_collect(const rm_bvh_node* __restrict__ nodes, const float* __restrict__ of, float dx, float dy, float dz,
                                        IvList& iv, int cap) {
    const double o[3] = {(double)of[0], (double)of[1], (double)of[2]};
    const float d[3] = {dx, dy, dz};
    int stack[kBvhStack];
    int sp = 0;
    stack[sp++] = 0;
    int n = 0;
    double invD[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) invD[i] = 1.0 / (double)d[i];
    const double tMin = 0.0, tMax = 10.0;
    while (sp > 0) {
        int ni = stack[--sp];
        const rm_bvh_node* nd = nodes + ni;
        double tEnter, tExit;
        if (!box_intersect_ray(nd->bmin, nd->bmax, o, d, invD, tEnter, tExit)) continue;
        if (tExit < tMin || tEnter > tMax) continue;
        double cEnter = tEnter > tMin ? tEnter : tMin;
        double cExit = tExit < tMax ? tExit : tMax;
        int left = nd->left, right = nd->right;
        if (left >= 0 || right >= 0) {
            if (left >= 0 && sp < kBvhStack) stack[sp++] = left;
            if (right >= 0 && sp < kBvhStack) stack[sp++] = right;  // popped first (bvh.ts:160-161)
        } else if (nd->prim_count > 0) {
            // stable insertion: after every element with enter <= cEnter (later DFS order sorts later)
            if (n == cap && !(cEnter < iv.enter[n - 1])) continue;
            int pos = (n < cap) ? n : cap - 1;
            while (pos > 0 && cEnter < iv.enter[pos - 1]) {
                iv.enter[pos] = iv.enter[pos - 1];
                iv.exit_[pos] = iv.exit_[pos - 1];
                --pos;
            }
            iv.enter[pos] = cEnter;
            iv.exit_[pos] = cExit;
            if (n < cap) ++n;
        }
    }
    return n;
}

// ------------------------------------------------------------------------------------------
// Fast path: the same BVH questions answered through the uniform leaf grid (rm_host.h LeafGrid).
// Every accepted leaf still goes through the reference's exact box test / slab test; the grid only
// decides which leaf boxes are looked at, with conservative cell ranges.
// ------------------------------------------------------------------------------------------
RM_DEV int grid_coord(float x, float o, float inv, int n) {
    int c = (int)floorf((x - o) * inv);
    return c < 0 ? 0 : (c >= n ? n - 1 : c);
}
RM_DEV bool cell_in_range(uint32_t lo, uint32_t hi, int x, int y, int z) {
    return x >= (int)(lo & 255u) && x <= (int)(hi & 255u) && y >= (int)((lo >> 8) & 255u) && y <= (int)((hi >> 8) & 255u) &&
           z >= (int)((lo >> 16) & 255u) && z <= (int)((hi >> 16) & 255u);
}

// Lazy, ordered generation of the per-ray leaf intervals of BVH.findRayIntersections (bvh.ts:126-178).
// A 3D-DDA walks the grid cells the ray crosses front to back.  A leaf is examined in the FIRST visited
// cell of its (convex) cell range, with the reference's exact slab test; accepted intervals wait in a small
// buffer sorted by (tEnter, right-first DFS order) and are released once the walk has passed their tEnter,
// so the consumer sees exactly the stably-sorted list of the reference — but only as far as the march needs.
constexpr int kPendCap = RM_PEND_CAP;  // rm_types.h
struct LazyIv {
    double tNext[3], tDelta[3], invD[3];  // invD = 1 / direction, computed once per ray
    float invDf[3];                       // (float)invD and whether the fp32 pre-test of the slab test applies to this ray
    bool f32ok;                           // (no axis-parallel component: |d_i| >= 1e-10 on every axis)
    double tEnd, safeT;
    double pEnter[kPendCap], pExit[kPendCap];
    int pNode[kPendCap];
    int cell[3], prev[3], step[3];
    int head, count;
    bool done, overflow;
};

RM_DEV void lazy_insert(LazyIv& lz, double enter, double exit_, int node) {
    if (lz.head + lz.count >= kPendCap) {  // compact
        for (int i = 0; i < lz.count; ++i) {
            lz.pEnter[i] = lz.pEnter[lz.head + i];
            lz.pExit[i] = lz.pExit[lz.head + i];
            lz.pNode[i] = lz.pNode[lz.head + i];
        }
        lz.head = 0;
        if (lz.count >= kPendCap) {
            // The buffer can no longer be the complete sorted prefix: drop it and hand over to the literal list (everything
            // popped so far was final, so the caller continues at the same cursor index).
            lz.overflow = true;
            lz.count = 0;
            return;
        }
    }
    int pos = lz.head + lz.count;
    // ascending tEnter; ties in right-first DFS order = descending pre-order node index (bvh.ts:141,160-161)
    while (pos > lz.head && (enter < lz.pEnter[pos - 1] || (enter == lz.pEnter[pos - 1] && node > lz.pNode[pos - 1]))) {
        lz.pEnter[pos] = lz.pEnter[pos - 1];
        lz.pExit[pos] = lz.pExit[pos - 1];
        lz.pNode[pos] = lz.pNode[pos - 1];
        --pos;
    }
    lz.pEnter[pos] = enter;
    lz.pExit[pos] = exit_;
    lz.pNode[pos] = node;
    lz.count++;
}

// Conservative fp32 form of BoundingBox.intersectRay + the [0, 10] clip for a ray with no axis-parallel component.  Every fp32
// slab parameter is within 2e-7 (relative) of the fp64 one the reference computes (f32 difference of two f32 values, f32
// reciprocal of the direction, one product: three roundings of 6e-8), so with m = 4e-7 x the largest |t| of the six planes the
// tests below only reject boxes the exact test rejects too; everything else goes on to the exact fp64 test.  NaN never rejects.
RM_DEV bool slab_certain_miss_f32(const float4 bmn, const float4 bmx, const float* __restrict__ of, const float (&invDf)[3]) {
    const float bl[3] = {bmn.x, bmn.y, bmn.z}, bh[3] = {bmx.x, bmx.y, bmx.z};
    float tMin = -3.0e38f, tMax = 3.0e38f, mag = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float t0 = (bl[i] - of[i]) * invDf[i], t1 = (bh[i] - of[i]) * invDf[i];
        tMin = fmaxf(tMin, fminf(t0, t1));
        tMax = fminf(tMax, fmaxf(t0, t1));
        mag = fmaxf(mag, fmaxf(fabsf(t0), fabsf(t1)));
    }
    const float m = 4.0e-7f * mag + 1.0e-30f;
    return (tMin - tMax > 2.f * m) || (tMax < -m) || (tMin > 10.f + m);
}

// returns false when the ray misses the root box (=> no intervals at all)
static __device__ __noinline__ bool lazy_init(const DevScene& sc, const float* __restrict__ of, float dx, float dy, float dz, LazyIv& lz) {
    const double o[3] = {(double)of[0], (double)of[1], (double)of[2]};
    const float d[3] = {dx, dy, dz};
    lz.head = 0;
    lz.count = 0;
    lz.done = false;
    lz.overflow = false;
    lz.safeT = -1.0;
    double invD[3];
    lz.f32ok = true;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        invD[i] = 1.0 / (double)d[i];
        lz.invD[i] = invD[i];
        lz.invDf[i] = (float)invD[i];
        lz.f32ok = lz.f32ok && !(fabs((double)d[i]) < 1e-10);
    }
    double tE, tX;
    const rm_bvh_node* root = sc.bvh;
    if (!box_intersect_ray(root->bmin, root->bmax, o, d, invD, tE, tX) || tX < 0.0 || tE > 10.0) {
        lz.done = true;
        return false;
    }
    const double t0 = tE > 0.0 ? tE : 0.0;
    lz.tEnd = tX < 10.0 ? tX : 10.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double p = o[k] + (double)d[k] * t0;
        const double org = (double)sc.grid_origin[k], cs = (double)sc.grid_cell[k];
        int c = (sc.grid_inv[k] > 0.f) ? (int)floor((p - org) * (double)sc.grid_inv[k]) : 0;
        c = c < 0 ? 0 : (c >= sc.grid_dims[k] ? sc.grid_dims[k] - 1 : c);
        lz.cell[k] = c;
        lz.prev[k] = -1;
        const double dk = (double)d[k];
        if (dk > 0.0 && cs > 0.0) {
            lz.step[k] = 1;
            lz.tDelta[k] = cs * invD[k];
            lz.tNext[k] = (org + (double)(c + 1) * cs - o[k]) * invD[k];
        } else if (dk < 0.0 && cs > 0.0) {
            lz.step[k] = -1;
            lz.tDelta[k] = -cs * invD[k];
            lz.tNext[k] = (org + (double)c * cs - o[k]) * invD[k];
        } else {
            lz.step[k] = 0;
            lz.tDelta[k] = d_inf();
            lz.tNext[k] = d_inf();
        }
    }
    return true;
}

// examine the current cell, then step to the next one
#if RM_INLINE_WALK
static __device__ __forceinline__
#else
static __device__ __noinline__
#endif
void lazy_advance_cell(const DevScene& sc, const float* __restrict__ of, float dx, float dy, float dz, LazyIv& lz) {
    const double o[3] = {(double)of[0], (double)of[1], (double)of[2]};
    const float d[3] = {dx, dy, dz};
    const double invD[3] = {lz.invD[0], lz.invD[1], lz.invD[2]};
    const int cx = lz.cell[0], cy = lz.cell[1], cz = lz.cell[2];
    const size_t c = ((size_t)cz * sc.grid_dims[1] + cy) * sc.grid_dims[0] + cx;
    // First cell of the walk: every leaf of the cell.  Later cells: only the leaves whose cell range starts here along the
    // step that brought us in (direction list lz.prev[0] = 2 * axis + (step < 0)) — the others were met in the previous cell.
    const int code = lz.prev[0];
    const uint32_t* __restrict__ list;
    uint32_t e0, e1;
    if (code < 0) {
        list = sc.grid_cell_node;
        e0 = __ldg(sc.grid_cell_start + c);
        e1 = __ldg(sc.grid_cell_start + c + 1);
    } else {
        const uint4 cd = __ldg(sc.grid_cell_dir + c);
        const unsigned cnt[6] = {cd.y & 0xFFFFu, cd.y >> 16, cd.z & 0xFFFFu, cd.z >> 16, cd.w & 0xFFFFu, cd.w >> 16};
        list = sc.grid_dir_node;
        e0 = cd.x;
#pragma unroll
        for (int k = 0; k < 5; ++k) e0 += (k < code) ? cnt[k] : 0u;
        unsigned len = cnt[0];
#pragma unroll
        for (int k = 1; k < 6; ++k) len = (k == code) ? cnt[k] : len;
        e1 = e0 + len;
    }
    const bool pre = lz.f32ok && RM_SLAB_PRETEST;
    const float invDf[3] = {lz.invDf[0], lz.invDf[1], lz.invDf[2]};
    const float4* __restrict__ lrec = reinterpret_cast<const float4*>(sc.grid_leafrec);
    for (uint32_t e = e0; e < e1; ++e) {
        const uint32_t node = __ldg(list + e);
        double tE, tX;
        if (lrec != nullptr) {
            // translation-only spheres: the box comes from the 80-byte leaf record the point query reads too (two LDG.128,
            // same cache lines) and first meets the fp32 pre-test: most leaves of a cell are missed by the ray
            const float4 bmn = __ldg(lrec + 5u * (size_t)node), bmx = __ldg(lrec + 5u * (size_t)node + 1);
            if (pre && slab_certain_miss_f32(bmn, bmx, of, invDf)) continue;
            const float bl[3] = {bmn.x, bmn.y, bmn.z}, bh[3] = {bmx.x, bmx.y, bmx.z};
            if (!box_intersect_ray(bl, bh, o, d, invD, tE, tX)) continue;
        } else {
            const rm_bvh_node* nd = sc.bvh + node;
            if (!box_intersect_ray(nd->bmin, nd->bmax, o, d, invD, tE, tX)) continue;
        }
        if (tX < 0.0 || tE > 10.0) continue;
        if (lz.overflow) break;
        lazy_insert(lz, tE > 0.0 ? tE : 0.0, tX < 10.0 ? tX : 10.0, (int)node);
    }
    if (lz.count > sc.lazy_cap) {  // RM_LAZY_CAP test knob (normally the buffer size): force the hand-over
        lz.overflow = true;
        lz.count = 0;
    }
    int ax = 0;
    if (lz.tNext[1] < lz.tNext[ax]) ax = 1;
    if (lz.tNext[2] < lz.tNext[ax]) ax = 2;
    const double tOut = lz.tNext[ax];
    lz.safeT = tOut - 1e-5;
    if (!(tOut < lz.tEnd)) {
        lz.done = true;
        return;
    }
    lz.prev[0] = 2 * ax + (lz.step[ax] < 0 ? 1 : 0);
    const int nc = lz.cell[ax] + lz.step[ax];
    if (nc < 0 || nc >= sc.grid_dims[ax]) {
        lz.done = true;
        return;
    }
    lz.cell[ax] = nc;
    lz.tNext[ax] += lz.tDelta[ax];
}

// next interval of the sorted list, or false when the list is exhausted
RM_DEV bool lazy_pop(const DevScene& sc, const float* __restrict__ of, float dx, float dy, float dz, LazyIv& lz, double& enter, double& exit_) {
    for (;;) {
        if (lz.count > 0 && (lz.done || lz.pEnter[lz.head] < lz.safeT)) {
            enter = lz.pEnter[lz.head];
            exit_ = lz.pExit[lz.head];
            lz.head++;
            lz.count--;
            if (lz.count == 0) lz.head = 0;
            return true;
        }
        if (lz.done || lz.overflow) return false;
        lazy_advance_cell(sc, of, dx, dy, dz, lz);
    }
}

// Move the interval cursor to the next entry of the sorted list (currentIntervalIdx++).  Lazy mode pulls it
// from the grid walk; if the walk's buffer overflowed, the ray falls back to the literal eager list.
template <class NP, bool kLazy>
RM_DEV void bvh_advance(const DevScene& sc, const float* __restrict__ of, Ray<NP>& r, IvList& iv, LazyIv& lz, int cap) {
    r.cur++;
    if constexpr (kLazy) {
        if (r.lazy) {
            double enter = 0.0, exit_ = 0.0;
            r.curValid = lazy_pop(sc, of, r.d[0], r.d[1], r.d[2], lz, enter, exit_);
            if (r.curValid) {
                r.curEnter = enter;
                r.curExit = exit_;
            }
            if (r.curValid || !lz.overflow) return;
            r.nIv = bvh_collect(sc.bvh, of, r.d[0], r.d[1], r.d[2], iv, cap);  // overflow: rebuild the list the reference's way
            r.lazy = false;
        }
    }
    r.curValid = r.cur < r.nIv;
    if (r.curValid) {
        r.curEnter = iv.enter[r.cur];
        r.curExit = iv.exit_[r.cur];
    }
}

// BVH.onRayMarchStep (bvh.ts:204-240).
template <class NP, bool kLazy>
RM_DEV double bvh_step(const DevScene& sc, const float* __restrict__ of, Ray<NP>& r, IvList& iv, LazyIv& lz, int cap) {
    if (!r.curValid) return -1.0;  // currentIntervalIdx >= intervals.length
    if (r.t < r.curEnter) return r.curEnter - r.t;
    if (r.t > r.curExit) {
        bvh_advance<NP, kLazy>(sc, of, r, iv, lz, cap);
        if (r.curValid) {
            if (r.curEnter > r.t) return r.curEnter - r.t;
        } else {
            return -1.0;
        }
    }
    return 0.0;
}

// Octree.findNode (octree.ts:223-248).  Children tile their parent exactly (shared f32 planes), and
// "first containing child wins" on shared faces means the low half whenever p <= centre.
RM_DEV int octree_find(const rm_octree_node* __restrict__ nodes, const float q[3]) {
    if (!box_contains(nodes[0].bmin, nodes[0].bmax, q)) return -1;
    int ni = 0;
    for (int lvl = 0; lvl <= kOctreeMaxDepth; ++lvl) {
        int fc = nodes[ni].first_child;
        if (fc < 0 || nodes[ni].level == kOctreeMaxDepth) return ni;
        const float* c = nodes[fc].bmax;  // child 0 = low octant; its max corner is the parent's centre
        int ci = (q[0] > c[0] ? 1 : 0) + (q[1] > c[1] ? 2 : 0) + (q[2] > c[2] ? 4 : 0);
        int child = fc + ci;
        if (!box_contains(nodes[child].bmin, nodes[child].bmax, q)) return ni;  // octree.ts:247
        ni = child;
    }
    return ni;
}

// Octree.marchRay (octree.ts:252-278) incl. intersectRayBox (:195-220: f32 tMin/tMax, no parallel guard,
// so +-Infinity and NaN flow through Math.max/min exactly as in JS).
RM_DEV double octree_march(const rm_octree_node* __restrict__ nodes, const double o[3], const float d[3], double t,
                           const float p[3]) {
    int ni = octree_find(nodes, p);
    if (ni < 0) return 0.0;
    const rm_octree_node* nd = nodes + ni;
    if (!nd->is_empty) return 0.0;
    double tMinV[3], tMaxV[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double invD = 1.0 / (double)d[i];
        double t0 = ((double)nd->bmin[i] - o[i]) * invD;
        double t1 = ((double)nd->bmax[i] - o[i]) * invD;
        if (invD < 0.0) {
            double tmp = t0;
            t0 = t1;
            t1 = tmp;
        }
        tMinV[i] = (double)f32r(t0);
        tMaxV[i] = (double)f32r(t1);
    }
    double tEnter = jsmax(jsmax(tMinV[0], tMinV[1]), tMinV[2]);
    double tExit = jsmin(jsmin(tMaxV[0], tMaxV[1]), tMaxV[2]);
    if (tEnter > tExit || tExit < 0.0) return 0.0;
    double toExit = jsmax(0.0, tExit - t);
    double step = jsmax(0.0, jsmin(toExit, nd->min_distance * 0.99));
    return step > 0.0 ? step + 0.001 : 0.0;
}

// ------------------------------------------------------------------------------------------
// Shaders (pure functions of the quantised per-pixel values)
// ------------------------------------------------------------------------------------------
RM_DEV uchar4 shade_heat(unsigned count_u16) {  // SDFHeatmap.ts:24-29 / IterationHeatmap.ts:24-29
    unsigned s = (count_u16 * 5u) % 256u;
    unsigned r = 2u * s;
    int g = 512 - 2 * (int)s;
    return make_uchar4((unsigned char)(r > 255u ? 255u : r), (unsigned char)(g > 255 ? 255 : g), 0, 255);
}
template <class NP>
RM_DEV uchar4 shade_phong(unsigned depth_u8, unsigned n0, unsigned n1, unsigned n2) {  // phongModel.ts:15-73
    typedef typename NP::F F;
    if (depth_u8 >= 255u) return make_uchar4(10, 10, 20, 255);
    // lightDir = normalize(vec3(1,-1,1.5))
    F ll = (F)1 * (F)1 + (F)-1 * (F)-1 + (F)1.5 * (F)1.5;
    F linv = NP::rsqrt_(ll);
    F L0 = NP::st((F)1 * linv), L1 = NP::st((F)-1 * linv), L2 = NP::st((F)1.5 * linv);
    F a0 = NP::st((F)n0 / (F)127.5 - (F)1.0);
    F a1 = NP::st((F)n1 / (F)127.5 - (F)1.0);
    F a2 = NP::st((F)n2 / (F)127.5 - (F)1.0);
    F len = a0 * a0 + a1 * a1 + a2 * a2;
    if (len > (F)0) len = NP::rsqrt_(len);
    F N0 = NP::st(a0 * len), N1 = NP::st(a1 * len), N2 = NP::st(a2 * len);
    F ndl = N0 * L0 + N1 * L1 + N2 * L2;
    F diffuse = ndl > (F)0 ? ndl : (F)0;  // Math.max(dot, 0); NaN cannot arise from u8 inputs
    F sc = (F)2 * ndl;
    F r0 = NP::st(N0 * sc), r1 = NP::st(N1 * sc), r2 = NP::st(N2 * sc);
    r0 = NP::st(r0 - L0);
    r1 = NP::st(r1 - L1);
    r2 = NP::st(r2 - L2);
    F rl = r0 * r0 + r1 * r1 + r2 * r2;
    if (rl > (F)0) rl = NP::rsqrt_(rl);
    r0 = NP::st(r0 * rl);
    r1 = NP::st(r1 * rl);
    r2 = NP::st(r2 * rl);
    F vdr = (F)0 * r0 + (F)0 * r1 + (F)1 * r2;  // dot(viewDir = (0,0,1), reflectDir)
    F specular = (F)0.5 * NP::pow32(vdr > (F)0 ? vdr : (F)0);
    F isum = (F)0.1 + diffuse + specular;
    F intensity = isum < (F)1 ? isum : (F)1;
    F depthFactor = (F)1 - (F)depth_u8 / (F)255;
    F color = (F)255 * intensity * depthFactor;
    unsigned char c = (unsigned char)to_u8_clamp(color);
    return make_uchar4(c, c, c, 255);
}
template <class NP>
RM_DEV uchar4 shade_pixel(int shader, unsigned depth_u8, unsigned n0, unsigned n1, unsigned n2, unsigned sdf_u16,
                          unsigned iters_u16) {
    switch (shader) {
        case RM_SHADER_PHONG: return shade_phong<NP>(depth_u8, n0, n1, n2);
        case RM_SHADER_SDF_HEATMAP: return shade_heat(sdf_u16);
        case RM_SHADER_ITERATION_HEATMAP: return shade_heat(iters_u16);
        default: return make_uchar4((unsigned char)n0, (unsigned char)n1, (unsigned char)n2, 255);  // normalModel.ts:21-24
    }
}

template <class NP>
__global__ void shade_kernel(ShadeParams P) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_pixels) return;
    uchar4 c = shade_pixel<NP>(P.shader, P.depth[i], P.normal[3 * (size_t)i], P.normal[3 * (size_t)i + 1],
                               P.normal[3 * (size_t)i + 2], P.sdf[i], P.iters[i]);
    reinterpret_cast<uchar4*>(P.rgba)[i] = c;
}

// ------------------------------------------------------------------------------------------
// The render kernel
// ------------------------------------------------------------------------------------------
struct LaneStats {
    unsigned long long sum_sdf = 0, sum_iters = 0, sum_sdf_full = 0, sum_iters_full = 0;
    unsigned long long ev_sphere = 0, ev_box = 0, ev_torus = 0, n_hit = 0, op_flops = 0;
    unsigned max_sdf = 0, min_sdf = 0xffffffffu, max_iters = 0, min_iters = 0xffffffffu;
};

RM_DEV unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

template <class NP, int ACCEL, int PK>
struct CtaShape {
    static constexpr int kWarps = (!NP::kExact && ACCEL == RM_ACCEL_BVH) ? (PK == PK_TSPHERE ? RM_TC_WARPS : 8) : 4;
};
// TCK: the instance whose all-primitives pass is the tensor-core cluster screen (translation-only spheres behind a BVH with the
//      cluster data uploaded); its FFMA search body is compiled out — scenes without cluster data (RM_DISABLE_TC, < 256 spheres)
//      launch the TCK = false instance instead.  113 KB less SASS in the hot kernel: -2.9 % frame time (profiles/r01_ab_tconly.log).
// ALGT: >= 0 fixes Job.algorithm at compile time (0 = sphere tracer, the BASELINE configs' algorithm): the V2 / V3 / fixed-step
//      state and branches fold away; -1 = run-time switch.
template <class NP, int ACCEL, int PK, bool TCK = false, int ALGT = -1>
__global__ void __launch_bounds__(32 * CtaShape<NP, ACCEL, PK>::kWarps,
                                  (ACCEL == RM_ACCEL_BVH) ? ((RM_MIN_BLOCKS_BVH * 4) / CtaShape<NP, ACCEL, PK>::kWarps > 0 ? (RM_MIN_BLOCKS_BVH * 4) / CtaShape<NP, ACCEL, PK>::kWarps : 1) : RM_MIN_BLOCKS_OTHER)
    render_kernel(const __grid_constant__ RenderParams P) {
    constexpr int kWarpsPerCta = CtaShape<NP, ACCEL, PK>::kWarps;
    constexpr int kQueueCap = kWarpsPerCta * 32;  // each thread has at most one request outstanding
    constexpr unsigned kBatch = (kWarpsPerCta >= 8) ? 64u : 32u;  // requests served per cooperative pass (FFMA search)
    // translation-only spheres behind a BVH: the pass runs on the tensor cores, 128 requests at a time (tc_pass)
    static_assert(!TCK || (!NP::kExact && ACCEL == RM_ACCEL_BVH && PK == PK_TSPHERE && kWarpsPerCta >= 16), "TCK needs the 16-warp sphere/BVH shape");
    constexpr bool kTC = TCK;
    constexpr unsigned kBatchMax = kTC ? (unsigned)kTcBlock : kBatch;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;

    const double o[3] = {(double)P.origin[0], (double)P.origin[1], (double)P.origin[2]};
    const double MAX_DIST = 10.0, EPSILON = 0.001;
    const int alg = (ALGT >= 0) ? ALGT : P.algorithm;
    const bool hitOnly = (alg == RM_ALG_FIXED_STEP || alg == RM_ALG_ADAPTIVE_STEP);  // return hit ? t : MAX_DIST
    const int maxSteps = hitOnly ? kMaxStepsFixed : kMaxStepsSphere;
    const double stepSize = P.step_size, overshoot = P.overshoot;
    const int bandH = P.y_end - P.y_start;

    Ray<NP> r;
    r.phase = PH_IDLE;
    r.cur = 0;
    r.nIv = 0;
    IvList iv;  // only touched when ACCEL == BVH (lives in local memory)
    LazyIv lz;  // fast-path BVH: grid walk state (local memory)
    constexpr bool kLazy = !NP::kExact && ACCEL == RM_ACCEL_BVH;
    LaneStats st;
    // Diagnostics accumulators.  Per-lane registers (13 values that stay live across the whole state machine) or — the default —
    // one set per warp in shared memory, updated with shared atomics when a pixel retires: the accumulators are touched once per
    // pixel, the registers they would pin are needed by the march on every iteration.
    // (measured, profiles/r02g_ab.jsonl: shared accumulators win on every kernel — cfg2 -25 %, cfg3 -17 %, cfg5 -5 % — except the
    // tensor-core instance, where per-lane registers are 1.5 % faster)
    constexpr bool kSmemStats = RM_SMEM_STATS != 0 && !TCK;
    __shared__ unsigned long long shStatSum[kSmemStats ? CtaShape<NP, ACCEL, PK>::kWarps : 1][9];
    __shared__ unsigned shStatMM[kSmemStats ? CtaShape<NP, ACCEL, PK>::kWarps : 1][4];
    if constexpr (kSmemStats) {
        if ((threadIdx.x & 31) < 9) shStatSum[threadIdx.x >> 5][threadIdx.x & 31] = 0ull;
        if ((threadIdx.x & 31) < 4) shStatMM[threadIdx.x >> 5][threadIdx.x & 31] = ((threadIdx.x & 1) ? 0xffffffffu : 0u);  // max, min, max, min
        __syncwarp();
    }

    // ---- shared memory: per-warp TMA stages for the primitive stream, and the CTA-wide request queue of
    //      the all-primitives service (requests are just a point, so they can move between warps even
    //      though ray state cannot: any warp that finds 32 of them serves them at full lane occupancy) ----
    // [2][kWarpsPerCta][kStageBytes / 16], stage-major and sized at launch: scenes small enough to stay resident in stage 0 get
    // only the first half (and the exact kernels none at all) — every KB of shared memory a CTA does not ask for is L1
    extern __shared__ __align__(128) float4 shStageDyn[];
    float4 (*shStage)[kWarpsPerCta][kStageBytes / 16] = reinterpret_cast<float4 (*)[kWarpsPerCta][kStageBytes / 16]>(shStageDyn);
    __shared__ __align__(8) unsigned long long shBar[kWarpsPerCta][2];
    __shared__ float4 shReq[kQueueCap];               // ring of requests: x, y, z, owner thread
    __shared__ double shRes[kWarpsPerCta * 32];        // results by owner thread
    __shared__ unsigned shReady[kWarpsPerCta * 32];
    __shared__ float shPartBest[kWarpsPerCta * kBatch];  // partial search results of the batch in flight: [warp][request] (FFMA) or [half][row] (TC)
    __shared__ int shPartCode[kWarpsPerCta * kBatch];
    __shared__ unsigned shTail, shHead, shGo, shStuck, shFinished;
    __shared__ unsigned shBandFin[kWarpsPerCta][kMaxBands];  // early download: pixels this warp finalised per row band, not yet published
    __shared__ __align__(8) unsigned long long shTcBar[kTcStages + kTcGroups];
    __shared__ unsigned shTcDrain[kTcGroups];
    __shared__ unsigned long long shTcKey[kTcBlock];  // per request: (flipped fp32 distance << 32 | sorted sphere index), atomicMin
    __shared__ unsigned shTcItems;
    __shared__ uint32_t shTmemBase;
    const int warpId = threadIdx.x >> 5;
    // Vectorised tile epilogue.  A retiring ray does not store its pixel: it keeps the quantised outputs packed in four registers
    // of its (now dead) ray state and waits as PH_HELD.  When the warp next wants to refill, the held pixels are written together:
    // if all 32 lanes hold the pixels of one whole 8x4 tile claimed in one go (lane = pixel index), the tile is transposed with
    // warp shuffles and written with 16-byte stores (u16 planes: one per tile row; RGBA: two per tile row) and 8-byte stores
    // (depth: one per tile row; normal: three per tile row); otherwise (edge tiles, tiles claimed piecewise while lanes were
    // parked on the all-primitives pass, unaligned planes) every held lane writes its own pixel as before.
    constexpr bool kHeldEpi = RM_VEC_EPILOGUE != 0;
    const bool vecOK = kHeldEpi && P.vec_store != 0;
    bool curWhole = false;  // warp-uniform: the current tile was claimed by all 32 lanes at once
    WarpStage ws;
    ws.phase = 0u;
#pragma unroll
    for (int sgi = 0; sgi < 2; ++sgi) {
        ws.buf[sgi] = shStage[sgi][warpId];
        ws.bufAddr[sgi] = smem_u32(shStage[sgi][warpId]);
        ws.bar[sgi] = smem_u32(&shBar[warpId][sgi]);
    }
    ws.resident = false;
    if constexpr (!NP::kExact && !(kTC && RM_TC_STATIC_QUEUE)) {  // (the tensor-core instance never streams primitives through the per-warp stages)
        if (lane == 0) {
            mbar_init(ws.bar[0], 1);
            mbar_init(ws.bar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        // small scenes: stage the whole primitive array once and keep it resident in shared memory
        constexpr int kPerStage0 = kStageBytes / (16 * ((PK == PK_TSPHERE) ? 1 : 4));
        if (P.scene.n_prims > 0 && ((PK == PK_TSPHERE) ? P.scene.n_chunks <= 4 : P.scene.n_prims <= kPerStage0)) {
            // ONE copy per CTA, in the first stage of the dynamic shared memory (the launcher allocates just that: 2.5 KB instead of
            // 20-80 KB, the difference is L1); warp 0 brings it in, the __syncthreads() below publishes it to the other warps
            const unsigned bytes = (PK == PK_TSPHERE) ? (unsigned)(P.scene.n_chunks * kChunkBytes) : (unsigned)(P.scene.n_prims * 64);
            ws.buf[0] = shStage[0][0];
            ws.bufAddr[0] = smem_u32(shStage[0][0]);
            if (warpId == 0) {
                if (lane == 0) {
                    mbar_expect_tx(ws.bar[0], bytes);
                    bulk_g2s(ws.bufAddr[0], P.scene.rec, bytes, ws.bar[0]);
                }
                mbar_wait(ws.bar[0], 0u);
                ws.phase ^= 1u;
            }
            ws.resident = true;  // (never streamed through again: the cooperative queue needs >= 256 primitives)
        }
    }
    shReady[threadIdx.x] = 0u;
    if (threadIdx.x < kWarpsPerCta * kMaxBands) (&shBandFin[0][0])[threadIdx.x] = 0u;
    if (threadIdx.x == 0) {
        shTail = 0u;
        shHead = 0u;
        shGo = 0u;
        shStuck = 0u;
        shFinished = 0u;
    }
    __syncthreads();
    bool wasStuck = false, wasFinished = false;  // this warp's contribution to shStuck / shFinished
    // the shared queue pays off when one all-primitives pass is much more expensive than a march step
    // (the tensor-core instance is only launched for >= 256 spheres with the cluster data uploaded: a compile-time `true`
    // drops the warp-local streamed search — a quarter of this instance's SASS — from the kernel)
    const bool useQueue = (kTC && RM_TC_STATIC_QUEUE) ? true : (!NP::kExact && (ACCEL == RM_ACCEL_BVH) && P.scene.n_prims >= 256);
    const bool useTC = kTC;
    const unsigned batchCap = useTC ? min(kBatchMax, (unsigned)RM_TC_BATCH) : kBatch;
    TcCtx tc;
    if constexpr (kTC) {
        if (useTC) {  // CTA-uniform
            // the dynamic shared memory (per-warp stages of the FFMA search) is re-used as A tile + B-tile ring
            static_assert((size_t)kWarpsPerCta * 2 * kStageBytes >= (size_t)kTcATileBytes + (size_t)kTcStages * kTcTileBytes + 4u * kTcItemCap, "TC tiles + item list do not fit");
            tc.sA = smem_u32(shStageDyn);
            tc.sB = tc.sA + kTcATileBytes;
            tc.barFull = smem_u32(&shTcBar[0]);
            tc.barTmemFull = smem_u32(&shTcBar[kTcStages]);
            tc.drainCnt = smem_u32(&shTcDrain[0]);
            tc.g = 0u;
            if (warpId == 0) {  // all 512 TMEM columns of the SM: one CTA per SM
                asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&shTmemBase)), "r"(512u) : "memory");
                asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
            }
            if (threadIdx.x == 0) {
                for (int i = 0; i < kTcStages; ++i) mbar_init(tc.barFull + 8u * i, 1);
                for (int i = 0; i < kTcGroups; ++i) {
                    mbar_init(tc.barTmemFull + 8u * i, 1);
                    shTcDrain[i] = 0u;
                }
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            tc.tmem = *(volatile uint32_t*)&shTmemBase;
            if (RM_RESIDENT_B && P.scene.n_tc_blocks <= kTcStages) {  // resident B tiles (see tc_pass): one TMA bulk copy per block, once
                if (threadIdx.x == 0) {
                    for (int b = 0; b < P.scene.n_tc_blocks; ++b) {
                        mbar_expect_tx(tc.barFull + 8u * (unsigned)b, kTcTileBytes);
                        bulk_g2s(tc.sB + (unsigned)b * kTcTileBytes, reinterpret_cast<const char*>(P.scene.tc_tiles) + (size_t)b * kTcTileBytes, kTcTileBytes,
                                 tc.barFull + 8u * (unsigned)b);
                    }
                }
                for (int b = 0; b < P.scene.n_tc_blocks; ++b) mbar_wait(tc.barFull + 8u * (unsigned)b, 0u);
            }
        }
    }

    if (P.anatomy && threadIdx.x == 0) atomicMin(&P.stats->t_enter_min, globaltimer_ns());
    // warp-uniform work-queue cursor
    int tile = -1, tilePos = kTileW * kTileH;
    unsigned tileT0 = 0u;  // clock() at the fetch of the current tile (cost-ordered queue)
    bool queueEmpty = false;

    // Early download (rm_render into page-locked planes): publish this warp's finished-pixel counts per row band;
    // the thread that completes a band raises its flag in host memory and the host starts that band's D2H while the
    // kernel is still running.  Called warp-uniformly once per tile and at exit: one fence per tile, not per pixel.
    auto publish_bands = [&]() {
        if (!P.band_flags) return;
        __syncwarp();  // orders every lane's pixel stores before lane b's fence below
        if (lane < kMaxBands) {
            const unsigned n = shBandFin[warpId][lane];
            if (n) {
                shBandFin[warpId][lane] = 0u;
                // release at gpu scope: the pixel stores are performed before the count becomes visible.  Not
                // __threadfence() + atomicAdd: that fence also invalidates the SM's L1 (CCTL.IVALL), once per tile
                unsigned prev;
                asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(prev) : "l"(&P.stats->band_done[lane]), "r"(n) : "memory");
                if (prev + n == P.band_px[lane]) {
                    __threadfence_system();
                    *(volatile unsigned int*)&P.band_flags[lane] = 1u;
                }
            }
        }
        __syncwarp();
    };

    // Write the held pixels (called warp-uniformly with the ballot of PH_HELD lanes); their lanes become PH_IDLE.
    auto flush_held = [&](unsigned heldM) {
        const bool held = r.phase == PH_HELD;
        const unsigned w0 = __float_as_uint(r.q[0]), w1 = __float_as_uint(r.q[1]), c1 = __float_as_uint(r.q[2]), c2 = __float_as_uint(r.h[0]);
        const size_t idx = held ? (size_t)r.py * P.width + r.px : 0;
        if (vecOK && curWhole && heldM == kFull) {
            // one whole tile, lane = pixel index (row = lane >> 3, column = lane & 7): transpose with shuffles, store vectors
            const bool q4 = (lane & 3) == 0, q8 = (lane & 7) == 0;
            if (P.rgba) {  // 4 pixels = 16 bytes per store, two per tile row
                const unsigned y = __shfl_down_sync(kFull, c1, 1), z = __shfl_down_sync(kFull, c1, 2), w = __shfl_down_sync(kFull, c1, 3);
                if (q4) *reinterpret_cast<uint4*>(P.rgba + 4 * idx) = make_uint4(c1, y, z, w);
            }
            if (P.rgba2) {
                const unsigned y = __shfl_down_sync(kFull, c2, 1), z = __shfl_down_sync(kFull, c2, 2), w = __shfl_down_sync(kFull, c2, 3);
                if (q4) *reinterpret_cast<uint4*>(P.rgba2 + 4 * idx) = make_uint4(c2, y, z, w);
            }
            {  // the two u16 planes: 8 pixels = 16 bytes = one store per tile row each
                const unsigned a = __shfl_down_sync(kFull, w1, 1);
                const unsigned sp = __byte_perm(w1, a, 0x5410), ip = __byte_perm(w1, a, 0x7632);  // (even lane, odd lane) pairs
                const unsigned s1 = __shfl_down_sync(kFull, sp, 2), s2 = __shfl_down_sync(kFull, sp, 4), s3 = __shfl_down_sync(kFull, sp, 6);
                const unsigned i1 = __shfl_down_sync(kFull, ip, 2), i2 = __shfl_down_sync(kFull, ip, 4), i3 = __shfl_down_sync(kFull, ip, 6);
                if (q8) {
                    *reinterpret_cast<uint4*>(reinterpret_cast<char*>(P.sdf) + 2 * idx) = make_uint4(sp, s1, s2, s3);
                    *reinterpret_cast<uint4*>(reinterpret_cast<char*>(P.iters) + 2 * idx) = make_uint4(ip, i1, i2, i3);
                }
            }
            {  // depth: 8 pixels = 8 bytes per tile row
                const unsigned d1 = __shfl_down_sync(kFull, w0, 1);
                const unsigned b2 = __byte_perm(w0, d1, 0x0040) & 0xffffu;  // bytes (own depth, next lane's depth)
                const unsigned b4 = b2 | (__shfl_down_sync(kFull, b2, 2) << 16);
                const unsigned b8 = __shfl_down_sync(kFull, b4, 4);
                if (q8) *reinterpret_cast<uint2*>(P.depth + idx) = make_uint2(b4, b8);
            }
            {  // normal: 3 bytes per pixel, 8 pixels = 24 bytes = three 8-byte stores per tile row
                const unsigned t = w0 >> 8;
                const unsigned t1 = __shfl_down_sync(kFull, t, 1), t2 = __shfl_down_sync(kFull, t, 2), t3 = __shfl_down_sync(kFull, t, 3);
                const unsigned W0 = t | (t1 << 24), W1 = (t1 >> 8) | (t2 << 16), W2 = (t2 >> 16) | (t3 << 8);  // 4 pixels = 12 bytes
                const unsigned X0 = __shfl_down_sync(kFull, W0, 4), X1 = __shfl_down_sync(kFull, W1, 4), X2 = __shfl_down_sync(kFull, W2, 4);
                if (q8) {
                    char* n = reinterpret_cast<char*>(P.normal) + 3 * idx;
                    *reinterpret_cast<uint2*>(n) = make_uint2(W0, W1);
                    *reinterpret_cast<uint2*>(n + 8) = make_uint2(W2, X0);
                    *reinterpret_cast<uint2*>(n + 16) = make_uint2(X1, X2);
                }
            }
            if (P.band_flags && q8) atomicAdd(&shBandFin[warpId][r.py / P.band_rows], (unsigned)kTileW);  // early download
        } else if (held) {
            P.depth[idx] = (uint8_t)(w0 & 0xffu);
            P.normal[3 * idx + 0] = (uint8_t)((w0 >> 8) & 0xffu);
            P.normal[3 * idx + 1] = (uint8_t)((w0 >> 16) & 0xffu);
            P.normal[3 * idx + 2] = (uint8_t)(w0 >> 24);
            P.sdf[idx] = (uint16_t)(w1 & 0xffffu);
            P.iters[idx] = (uint16_t)(w1 >> 16);
            if (P.rgba) reinterpret_cast<unsigned*>(P.rgba)[idx] = c1;
            if (P.rgba2) reinterpret_cast<unsigned*>(P.rgba2)[idx] = c2;
            if (P.band_flags) atomicAdd(&shBandFin[warpId][r.py / P.band_rows], 1u);
        }
        if (held) r.phase = PH_IDLE;
    };

    // Warp scheduler.  Expensive, warp-serialising stages are deferred until enough lanes want them: BVH
    // ray set-up (a full tree traversal) runs when initLanes lanes are free or nothing else can progress;
    // the all-primitives pass goes through the CTA-wide request queue below.
    const int initLanes = RM_INIT_LANES;
    r.pending = false;

#ifdef RM_PHASE_TIMING
    long long tKernel0 = clock64(), tSearch = 0, tBarrier = 0, tStuck = 0;
#define RM_T0() long long t0_ = clock64()
#define RM_T1(acc) acc += clock64() - t0_
#else
#define RM_T0()
#define RM_T1(acc)
#endif
    for (;;) {
        // ---- (0) CTA rendezvous for the all-primitives pass (cooperative mode only) ----
        if (useQueue) {
            bool leave = false;
            for (;;) {
                unsigned go = 0;
                if (lane == 0) go = *(volatile unsigned*)&shGo;
                go = __shfl_sync(kFull, go, 0);
                if (go) {
                    // Every warp of the CTA joins: request i <-> lane i in each warp; warp w searches the stages
                    // c = w (mod W) of the primitive stream, partial minima are combined through shared memory.
#ifdef RM_PHASE_TIMING
                    long long tb0_ = clock64();
#endif
                    __syncthreads();
                    const unsigned head = *(volatile unsigned*)&shHead, tail = *(volatile unsigned*)&shTail;
#ifdef RM_PHASE_TIMING
                    tBarrier += (clock64() - tb0_) + (long long)(head & 0u);  // the volatile read forces the deferred barrier to resolve first
#endif
                    const unsigned nBatch = min(batchCap, tail - head);
                    bool tcDone = false;
                    if constexpr (kTC) {
                        if (useTC) {
                            RM_T0();
                            tc_pass(P, tc, shReq, head, nBatch, kQueueCap, shPartBest, shPartCode, shTcKey, &shTcItems);
                            RM_T1(tSearch);
                            tcDone = true;
                        }
                    }
                    // lane i takes request i (and request i + 32 in the 64-request form: two points per lane
                    // halve the shared-memory wavefronts per evaluation)
                    constexpr int NQ = (int)(kBatch / 32u);
                    float rq[NQ][3];
                    if constexpr (!kTC) {  // (the tensor-core instance carries no FFMA search body)
#pragma unroll
                    for (int k = 0; k < NQ; ++k) {
                        rq[k][0] = rq[k][1] = rq[k][2] = 0.f;
                        const unsigned ri = (unsigned)lane + 32u * (unsigned)k;
                        if (ri < nBatch) {
                            const float4 v = shReq[(head + ri) % kQueueCap];
                            rq[k][0] = v.x;
                            rq[k][1] = v.y;
                            rq[k][2] = v.z;
                        }
                    }
                    float pbest[NQ];
                    int pcode[NQ];
                    if constexpr (!NP::kExact) {
                        RM_T0();
                        if constexpr (PK == PK_TSPHERE) {
                            search_stages_ts<NQ>(P, rq, ws, lane, warpId, kWarpsPerCta, pbest, pcode);
                        } else {
#pragma unroll
                            for (int k = 0; k < NQ; ++k) search_stages<PK>(P, rq[k], ws, lane, warpId, kWarpsPerCta, pbest[k], pcode[k]);
                        }
                        RM_T1(tSearch);
                    } else {
#pragma unroll
                        for (int k = 0; k < NQ; ++k) {
                            pbest[k] = 10.f;
                            pcode[k] = -1;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < NQ; ++k) {
                        shPartBest[warpId * (int)kBatch + lane + 32 * k] = pbest[k];
                        shPartCode[warpId * (int)kBatch + lane + 32 * k] = pcode[k];
                    }
                    }  // !tcDone
#ifdef RM_PHASE_TIMING
                    long long tb1_ = clock64();
#endif
                    __syncthreads();
#ifdef RM_PHASE_TIMING
                    tBarrier += (clock64() - tb1_) + (long long)(*(volatile unsigned*)&shHead & 0u);
#endif
                    // combine: request i is finished by thread i (warp i / 32 ... only the first 32 threads have one)
                    if (threadIdx.x < nBatch) {
                        const float4 v = shReq[(head + threadIdx.x) % kQueueCap];
                        const float q3[3] = {v.x, v.y, v.z};
                        const int owner = __float_as_int(v.w);
                        double res;
                        if constexpr (!NP::kExact) {
                            float bb = shPartBest[threadIdx.x];
                            int bc = shPartCode[threadIdx.x];
                            const int nPart = tcDone ? kTcGroups : kWarpsPerCta, pStride = tcDone ? kTcBlock : (int)kBatch;
                            for (int w = 1; w < nPart; ++w) {
                                const float ob = shPartBest[w * pStride + (int)threadIdx.x];
                                if (ob < bb) {
                                    bb = ob;
                                    bc = shPartCode[w * pStride + (int)threadIdx.x];
                                }
                            }
                            res = finish_search<PK, kTC>(P, q3, bc);
                        } else {
                            res = 10.0;
                            for (int j = 0; j < P.scene.n_prims; ++j)
                                res = jsmin(prim_sdf_exact(P.scene, j, (double)q3[0], (double)q3[1], (double)q3[2], P.length_sqrt), res);
                        }
                        shRes[owner] = res;
                        shReady[owner] = 1u;
                    }
                    if (threadIdx.x == 0) {
                        shHead = head + nBatch;
                        shGo = (tail - head - nBatch >= batchCap) ? 1u : 0u;
                    }
                    __syncthreads();
                    continue;
                }
                if (!wasStuck) break;  // run a normal iteration
                // stuck: leave the wait as soon as one of our parked lanes has its result, or everybody is finished
                const unsigned got = __ballot_sync(kFull, r.pending && ((volatile unsigned*)shReady)[threadIdx.x] != 0u);
                if (got) break;
                unsigned fin = 0;
                if (lane == 0) fin = *(volatile unsigned*)&shFinished;
                fin = __shfl_sync(kFull, fin, 0);
                if (wasFinished && fin == (unsigned)kWarpsPerCta) {
                    leave = true;
                    break;
                }
                {
                    RM_T0();
                    __nanosleep(128);
                    RM_T1(tStuck);
                }
            }
            if (leave) break;
        }

        // ---- (a) refill: lanes without a ray claim the next pixels of the warp's current tile ----
        unsigned idle;
        {
            const unsigned heldM = kHeldEpi ? __ballot_sync(kFull, r.phase == PH_HELD) : 0u;
            const int nIdle = __popc(__ballot_sync(kFull, r.phase == PH_IDLE) | heldM);
            const int nPend = __popc(__ballot_sync(kFull, r.pending));
            const int nAct = 32 - nIdle - nPend;
            const bool doRefill = (nAct == 0) || (nIdle >= initLanes);
            if (kHeldEpi && doRefill && heldM) flush_held(heldM);  // held pixels are written before their lanes take new ones
            idle = doRefill ? __ballot_sync(kFull, r.phase == PH_IDLE) : 0u;
        }
        while (idle && !queueEmpty) {
            if (tilePos >= kTileW * kTileH) {
                publish_bands();
                unsigned t = 0;
                if (lane == 0) t = atomicAdd(&P.stats->queue, 1u);
                t = __shfl_sync(kFull, t, 0);
                if ((int)t >= P.n_tiles) {
                    queueEmpty = true;
                    if (P.anatomy && lane == 0) atomicMin(&P.stats->t_drain_min, globaltimer_ns());
                    break;
                }
                // cost-ordered queue: ticket t -> the t-th most expensive tile of the previous frame of this geometry (so the frame ends
                // on cheap tiles and the end-of-frame tail shrinks); every warp records what its tiles cost for the next frame
                if (P.tile_cost) {
                    const unsigned now = (unsigned)clock();
                    if (tile >= 0 && lane == 0) P.tile_cost[tile] = now - tileT0;
                    tileT0 = now;
                }
                tile = P.tile_order ? (int)__ldg(P.tile_order + t) : (int)t;
                tilePos = 0;
            }
            int avail = kTileW * kTileH - tilePos;
            int nIdle = __popc(idle);
            int take = avail < nIdle ? avail : nIdle;
            if (tilePos == 0) curWhole = (take == kTileW * kTileH);  // all 32 lanes at once: lane = pixel index within the tile
            int rank = __popc(idle & lt_mask);
            if (r.phase == PH_IDLE && rank < take) {
                int k = tilePos + rank;
                // tile -> (owned stripe, tile row inside the stripe, tile column); stripes interleave across GPUs
                int os = tile / P.tiles_per_stripe, rem = tile - os * P.tiles_per_stripe;
                int tyIn = rem / P.tiles_x, tx = rem - tyIn * P.tiles_x;
                int x = tx * kTileW + (k % kTileW);
                int yl = (os * P.stripe_count + P.stripe_index) * P.stripe_rows + tyIn * kTileH + (k / kTileW);
                if (x < P.width && yl < bandH) {  // edge tiles: out-of-range pixels are skipped
                    r.px = x;
                    r.py = yl;
                    r.phase = PH_NEW;
                }
            }
            tilePos += take;
            idle = __ballot_sync(kFull, r.phase == PH_IDLE);
        }
        if (!useQueue && queueEmpty && __ballot_sync(kFull, r.phase != PH_IDLE) == 0u) break;  // queue drained and every ray retired

        // ---- (b) ray set-up (raymarcher.ts:73-88 + onRayMarchStart) ----
        if (r.phase == PH_NEW) {
            int y = P.y_start + r.py;
            double v = ((double)y / (double)P.height - 0.5) * 2.0;
            double u = ((double)r.px / (double)P.width - 0.5) * 2.0;
            double a0 = (double)f32r(u), a1 = (double)f32r(v), a2 = -1.0;
            double b0 = (double)f32r(a0 * (double)P.rot3[0] + a1 * (double)P.rot3[3] + a2 * (double)P.rot3[6]);
            double b1 = (double)f32r(a0 * (double)P.rot3[1] + a1 * (double)P.rot3[4] + a2 * (double)P.rot3[7]);
            double b2 = (double)f32r(a0 * (double)P.rot3[2] + a1 * (double)P.rot3[5] + a2 * (double)P.rot3[8]);
            double len = b0 * b0 + b1 * b1 + b2 * b2;
            if (len > 0.0) len = 1.0 / sqrt(len);
            r.d[0] = f32r(b0 * len);
            r.d[1] = f32r(b1 * len);
            r.d[2] = f32r(b2 * len);
            r.t = 0.0;
            r.prevSDF = 0.0;
            r.prevStep = 0.0;
            r.sdf = 0;
            r.iters = 0;
            r.nSphere = 0;
            r.nBox = 0;
            if constexpr (NP::kExact) {
                r.ex.nTorus = 0;
                r.ex.opFlops = 0;
                r.ex.forceExact = false;
            }
            r.i = 0;
            r.done = false;
            r.cur = 0;
            r.depth = 0.0;
            r.phase = PH_STEP;
            if constexpr (ACCEL == RM_ACCEL_BVH) {
                r.cur = -1;
                r.nIv = 0;
                r.lazy = kLazy;
                if constexpr (kLazy) {
                    if (P.scene.grid_cell_dir == nullptr) {  // no direction lists (counts beyond 16 bits): the literal list
                        r.lazy = false;
                        r.nIv = bvh_collect(P.scene.bvh, P.origin, r.d[0], r.d[1], r.d[2], iv, maxSteps + 1);
                    } else if (!lazy_init(P.scene, P.origin, r.d[0], r.d[1], r.d[2], lz)) {
                        r.lazy = true;  // missed the root box: lz.done, empty list
                    }
                } else {
                    r.nIv = bvh_collect(P.scene.bvh, P.origin, r.d[0], r.d[1], r.d[2], iv, maxSteps + 1);
                }
                bvh_advance<NP, kLazy>(P.scene, P.origin, r, iv, lz, maxSteps + 1);  // cursor -> interval 0
                if (!r.curValid) {  // {terminate:true} -> return MAX_DIST (sphereTracer.ts:38-40)
                    r.depth = MAX_DIST;
                    r.done = true;
                }
            }
        }

        // ---- (c) top of the march loop: accel callbacks, skips, next query point ----
        if (r.phase == PH_STEP) {
            while (!r.done) {
                if (r.i >= maxSteps) {  // loop ran out of indices
                    r.depth = hitOnly ? MAX_DIST : r.t;
                    r.done = true;
                    break;
                }
                float p[3] = {f32r(o[0] + (double)r.d[0] * r.t), f32r(o[1] + (double)r.d[1] * r.t),
                              f32r(o[2] + (double)r.d[2] * r.t)};
                if constexpr (ACCEL != RM_ACCEL_NONE) {
                    double skip;
                    if constexpr (ACCEL == RM_ACCEL_BVH) skip = bvh_step<NP, kLazy>(P.scene, P.origin, r, iv, lz, maxSteps + 1);
                    else skip = octree_march(P.scene.oct, o, r.d, r.t, p);
                    if (skip == -1.0) {  // nothing left: `return MAX_DIST`
                        r.depth = MAX_DIST;
                        r.done = true;
                        break;
                    } else if (skip > 0.0) {
                        r.t += skip;
                        if (r.t > MAX_DIST) {  // `break` out of the for loop
                            r.depth = hitOnly ? MAX_DIST : r.t;
                            r.done = true;
                            break;
                        }
                        r.prevSDF = 0.0;  // adaptiveStepV2.ts:75-76 (unused by the other loops)
                        r.prevStep = 0.0;
                        r.i++;
                        continue;
                    }
                }
                r.q[0] = p[0];
                r.q[1] = p[1];
                r.q[2] = p[2];
                r.phase = PH_WAIT_MARCH;
                break;
            }
            if (r.phase == PH_STEP) {
                // march finished with r.depth: hit position + normal taps (raymarcher.ts:94-102)
                r.h[0] = f32r(o[0] + (double)r.d[0] * r.depth);
                r.h[1] = f32r(o[1] + (double)r.d[1] * r.depth);
                r.h[2] = f32r(o[2] + (double)r.d[2] * r.depth);
                if (r.depth >= MAX_DIST) {
                    r.n0 = 0.f;
                    r.n1 = 0.f;
                    r.n2 = 0.f;
                    r.phase = PH_FINAL;
                } else {
                    r.q[0] = r.h[0];
                    r.q[1] = r.h[1];
                    r.q[2] = r.h[2];
                    r.phase = PH_WAIT_N0;
                }
            }
        }

        // ---- (d) resolve the pending scene-distance query (scene.ts:144-190) ----
        const bool waiting = (r.phase >= PH_WAIT_MARCH && r.phase <= PH_WAIT_N3);
        unsigned svSphere = 0u, svBox = 0u, svTorus = 0u, svOp = 0u;  // exact kernels: counters before this query (fp32 object mode may repeat it)
        if constexpr (NP::kExact) {
            r.ex.fastq = P.fast_objects != 0 && !r.ex.forceExact && (r.phase == PH_WAIT_MARCH || r.phase == PH_WAIT_V3B);
            svSphere = r.nSphere;
            svBox = r.nBox;
            svTorus = r.ex.nTorus;
            svOp = r.ex.opFlops;
        }
        double dd = 10.0;      // the query result handed to the control logic
        float distF = 10.f;    // fast model: running fp32 min over the candidate primitives
        int argmin = -1;       // fast model: its primitive
        int argRec = -1;       // fast model, leaf records: 2 * node + slot of the winner when it is held inline
        unsigned cnt = 0;
        bool needAll = false, polish = false;
        if (waiting && !r.pending) {
            if constexpr (ACCEL == RM_ACCEL_NONE) {
                needAll = true;
            } else if constexpr (ACCEL == RM_ACCEL_OCTREE) {
                int ni = octree_find(P.scene.oct, r.q);
                if (ni < 0) {
                    needAll = true;  // outside the octree bounds: full evaluation (scene.ts:166)
                } else {
                    const rm_octree_node* nd = P.scene.oct + ni;
                    int pc = nd->prim_count;
                    if (pc > 0) {
                        leaf_prims<NP, PK>(P, P.scene.leaf_prims + nd->prim_first, pc, r.q, dd, distF, argmin, r);
                        cnt = (unsigned)pc;
                        polish = true;
                    } else if (nd->is_empty) {
                        dd = jsmin(10.0, nd->min_distance * 0.99);
                    }
                }
            } else {  // BVH.getPrimitivesAt (bvh.ts:95-121): every leaf whose box contains p, left before right
                const rm_bvh_node* nodes = P.scene.bvh;
                if constexpr (kLazy) {
                    // fast path: the leaves overlapping p's grid cell, each with the reference's box test
                    if (box_contains(nodes[0].bmin, nodes[0].bmax, r.q)) {
                        const int gx = grid_coord(r.q[0], P.scene.grid_origin[0], P.scene.grid_inv[0], P.scene.grid_dims[0]);
                        const int gy = grid_coord(r.q[1], P.scene.grid_origin[1], P.scene.grid_inv[1], P.scene.grid_dims[1]);
                        const int gz = grid_coord(r.q[2], P.scene.grid_origin[2], P.scene.grid_inv[2], P.scene.grid_dims[2]);
                        const size_t c = ((size_t)gz * P.scene.grid_dims[1] + gy) * P.scene.grid_dims[0] + gx;
                        const uint32_t e0 = P.scene.grid_cell_start[c], e1 = P.scene.grid_cell_start[c + 1];
                        for (uint32_t e = e0; e < e1; ++e) {
                            const uint32_t node = __ldg(P.scene.grid_cell_node + e);
                            if constexpr (PK == PK_TSPHERE) {
                                // one 80-byte record per leaf: box + spheres inline
                                const float4* lr = reinterpret_cast<const float4*>(P.scene.grid_leafrec + node);
                                const float4 a = __ldg(lr), b = __ldg(lr + 1);
                                if (!(r.q[0] >= a.x && r.q[0] <= b.x && r.q[1] >= a.y && r.q[1] <= b.y && r.q[2] >= a.z && r.q[2] <= b.z)) continue;
                                const int pc = __float_as_int(a.w);
                                if (pc > 0) {
#pragma unroll
                                    for (int k = 0; k < 2; ++k) {
                                        if (k < pc) {
                                            const float4 sp = __ldg(lr + 2 + k);
                                            const float lx = r.q[0] + sp.x, ly = r.q[1] + sp.y, lz2 = r.q[2] + sp.z;
                                            const float d = NumFast::sqrt_(fmaf(lx, lx, fmaf(ly, ly, lz2 * lz2))) - sp.w;
                                            if (d < distF) {
                                                distF = d;
                                                argRec = (int)(2u * node) + k;
                                                argmin = -1;
                                            }
                                        }
                                    }
                                    r.nSphere += (unsigned)pc;
                                    cnt += (unsigned)pc;
                                } else {  // a leaf with more than two primitives (depth limit of the builder): the generic path
                                    const int before = argmin;
                                    leaf_prims<NP, PK>(P, P.scene.leaf_prims + __float_as_int(b.w), -pc, r.q, dd, distF, argmin, r);
                                    if (argmin != before) argRec = -1;
                                    cnt += (unsigned)(-pc);
                                }
                            } else {
                                const rm_bvh_node* nd = nodes + node;
                                if (!box_contains(nd->bmin, nd->bmax, r.q)) continue;
                                const int pc = nd->prim_count;
                                leaf_prims<NP, PK>(P, P.scene.leaf_prims + nd->prim_first, pc, r.q, dd, distF, argmin, r);
                                cnt += (unsigned)pc;
                            }
                        }
                    }
                } else {
                    int stack[kBvhStack];
                    int sp = 0;
                    stack[sp++] = 0;
                    while (sp > 0) {
                        int ni = stack[--sp];
                        const rm_bvh_node* nd = nodes + ni;
                        if (!box_contains(nd->bmin, nd->bmax, r.q)) continue;
                        int left = nd->left, right = nd->right;
                        if (left < 0 && right < 0) {
                            int pc = nd->prim_count;
                            leaf_prims<NP, PK>(P, P.scene.leaf_prims + nd->prim_first, pc, r.q, dd, distF, argmin, r);
                            cnt += (unsigned)pc;
                        } else {
                            if (right >= 0 && sp < kBvhStack) stack[sp++] = right;  // popped after left
                            if (left >= 0 && sp < kBvhStack) stack[sp++] = left;
                        }
                    }
                }
                if (cnt == 0) needAll = true;  // candidates.length === 0 -> every primitive (scene.ts:173)
                else polish = true;
            }
        }
        if constexpr (!NP::kExact) {
            // fast model: one fp64 evaluation of the nearest candidate found by the fp32 search
            if (polish && argmin >= 0) {
                dd = jsmin(prim_sdf_polish<PK>(P.scene, argmin, r.q), 10.0);
            } else if (polish && argRec >= 0) {  // winner held inline in a leaf record: same arithmetic as prim_sdf_polish
                const LeafRecTS* lr = P.scene.grid_leafrec + (argRec >> 1);
                const float4 sp = __ldg(&lr->s[argRec & 1]);
                const double lx = (double)__fadd_rn(r.q[0], sp.x), ly = (double)__fadd_rn(r.q[1], sp.y), lz2 = (double)__fadd_rn(r.q[2], sp.z);
                dd = jsmin(sqrt(lx * lx + ly * ly + lz2 * lz2) - __ldg(&lr->r[argRec & 1]), 10.0);
            }
        }
        // ---- (d2) the dense all-primitives pass ----
        if (!useQueue) {
            // small scenes / no acceleration structure: the warp serves its own lanes right away
            const unsigned need = __ballot_sync(kFull, needAll);
            if (need) {
                bool fq = false;
                if constexpr (NP::kExact) fq = r.ex.fastq;
                double v = scene_all_prims<NP, PK>(P, r.q, ws, needAll, lane, fq);
                if (needAll) {
                    dd = v;
                    cnt = (unsigned)P.scene.n_prims;
                    r.nSphere += P.scene.type_hist[0];
                    r.nBox += P.scene.type_hist[1];
                    if constexpr (NP::kExact) {
                        r.ex.nTorus += P.scene.type_hist[2];
                        r.ex.opFlops += P.scene.all_op_flops;
                    }
                }
            }
        } else {
            // CTA-cooperative mode: requests go to the shared ring; the pass itself runs in bf_phase (loop top)
            volatile unsigned* vReady = shReady;
            // 1. owners pick up finished requests
            if (r.pending && vReady[threadIdx.x] != 0u) {
                dd = ((volatile double*)shRes)[threadIdx.x];
                vReady[threadIdx.x] = 0u;
                r.pending = false;
                cnt = (unsigned)P.scene.n_prims;
                r.nSphere += P.scene.type_hist[0];
                r.nBox += P.scene.type_hist[1];
            }
            // 2. lanes that need every primitive publish a request (warp-aggregated slot reservation)
            const unsigned need = __ballot_sync(kFull, needAll);
            if (need) {
                const int leader = __ffs(need) - 1;
                unsigned base = 0;
                if (lane == leader) base = atomicAdd(&shTail, (unsigned)__popc(need));
                base = __shfl_sync(kFull, base, leader);
                if (needAll) {
                    const unsigned slot = base + (unsigned)__popc(need & lt_mask);
                    shReq[slot % kQueueCap] = make_float4(r.q[0], r.q[1], r.q[2], __int_as_float((int)threadIdx.x));
                    r.pending = true;
                }
                __syncwarp();
                if (lane == leader) {
                    __threadfence_block();
                    if (base + (unsigned)__popc(need) - *(volatile unsigned*)&shHead >= batchCap) *(volatile unsigned*)&shGo = 1u;  // a full batch is waiting
                }
            }
        }

        // ---- (e) consume the query result ----
        bool repeatExact = false;
        if constexpr (NP::kExact) {
            // fp32 object mode: a distance within 2e-5 of the hit threshold decides nothing — the same query runs again, exactly,
            // in the next trip through the loop (its evaluations are not counted twice)
            if (waiting && r.ex.fastq && fabs(dd - EPSILON) < 2.0e-5) {
                repeatExact = true;
                r.ex.forceExact = true;
                r.nSphere = svSphere;
                r.nBox = svBox;
                r.ex.nTorus = svTorus;
                r.ex.opFlops = svOp;
            } else if (waiting) {
                r.ex.forceExact = false;
            }
        }
        if (waiting && !r.pending && !repeatExact) {
            r.sdf += cnt;
            switch (r.phase) {
                case PH_WAIT_MARCH: {
                    r.phase = PH_STEP;
                    r.iters++;
                    if (alg == RM_ALG_SPHERE_TRACER) {  // sphereTracer.ts:67-74
                        r.t += dd;
                        if (dd < EPSILON || r.t > MAX_DIST) {
                            r.depth = r.t;
                            r.done = true;
                        } else {
                            r.i++;
                        }
                    } else if (hitOnly) {  // fixedStep.ts:75-89 / adaptiveStep.ts:75-96
                        if (dd < EPSILON) {
                            r.depth = r.t;
                            r.done = true;
                        } else {
                            double step;
                            if (alg == RM_ALG_FIXED_STEP) {
                                step = stepSize;
                            } else if (dd < 0.1) {  // NEAR_DIST
                                step = 0.01;        // NEAR_STEP
                            } else {
                                step = 0.8 * dd;  // STEP_SCALE
                                const double MIN_STEP = 0.1 * 0.25, MAX_STEP = 0.1 * 5.0;
                                if (step < MIN_STEP) step = MIN_STEP;
                                if (step > MAX_STEP) step = MAX_STEP;
                            }
                            r.t += step;
                            if (r.t > MAX_DIST) {
                                r.depth = MAX_DIST;
                                r.done = true;
                            } else {
                                r.i++;
                            }
                        }
                    } else {  // V2 (adaptiveStepV2.ts:81-116) and V3 (adaptiveStepV3.ts:72-130)
                        if (dd < EPSILON || r.t > MAX_DIST) {
                            r.depth = r.t;
                            r.done = true;
                        } else if (r.i == 0 || r.prevSDF == 0.0) {
                            r.t += dd;
                            r.prevSDF = dd;
                            r.prevStep = dd;
                            r.i++;
                        } else if (r.prevStep <= (r.prevSDF + dd)) {
                            double step = dd * overshoot;
                            r.t += step;
                            r.prevSDF = dd;
                            r.prevStep = step;
                            r.i++;
                        } else if (alg == RM_ALG_ADAPTIVE_STEP_V2) {
                            r.t -= r.prevStep;
                            r.t += r.prevSDF;
                            r.prevStep = r.prevSDF;
                            r.i++;
                        } else {  // V3: step back, take the bridging tap d3
                            double originalPos = r.t - r.prevStep;
                            r.t = originalPos + r.prevSDF;
                            r.aux0 = originalPos;
                            r.aux1 = dd;  // newSDF
                            r.q[0] = f32r(o[0] + (double)r.d[0] * r.t);
                            r.q[1] = f32r(o[1] + (double)r.d[1] * r.t);
                            r.q[2] = f32r(o[2] + (double)r.d[2] * r.t);
                            r.phase = PH_WAIT_V3B;
                        }
                    }
                    break;
                }
                case PH_WAIT_V3B: {  // adaptiveStepV3.ts:113-134
                    r.iters++;
                    double newSDF = r.aux1, d3 = dd;
                    if (r.prevSDF + newSDF + d3 >= r.prevStep) {
                        r.t = r.aux0 + r.prevStep + newSDF;
                        r.prevSDF = newSDF;
                        r.prevStep = newSDF;
                    } else {
                        r.prevSDF = d3;
                        r.prevStep = d3;
                        r.t += d3;
                    }
                    r.i++;
                    r.phase = PH_STEP;
                    break;
                }
                case PH_WAIT_N0:  // raymarcher.ts:124-128
                    r.nd = dd;
                    r.q[0] = f32r((double)r.h[0] - 0.01);
                    r.phase = PH_WAIT_N1;
                    break;
                case PH_WAIT_N1:
                    r.n0 = f32r(r.nd - dd);  // n[0] = d - d(p - e_x): Float32Array store
                    r.q[0] = r.h[0];
                    r.q[1] = f32r((double)r.h[1] - 0.01);
                    r.phase = PH_WAIT_N2;
                    break;
                case PH_WAIT_N2:
                    r.n1 = f32r(r.nd - dd);
                    r.q[1] = r.h[1];
                    r.q[2] = f32r((double)r.h[2] - 0.01);
                    r.phase = PH_WAIT_N3;
                    break;
                default: {  // PH_WAIT_N3: normalise (raymarcher.ts:131-133)
                    r.n2 = f32r(r.nd - dd);
                    double x = (double)r.n0, y = (double)r.n1, z = (double)r.n2;
                    double len = x * x + y * y + z * z;
                    if (len > 0.0) len = 1.0 / sqrt(len);
                    r.n0 = f32r(x * len);
                    r.n1 = f32r(y * len);
                    r.n2 = f32r(z * len);
                    r.phase = PH_FINAL;
                    break;
                }
            }
        }

        // ---- (f) finalize: quantise, shade, store, accumulate diagnostics (raymarcher.ts:103-106) ----
        if (r.phase == PH_FINAL) {
            size_t idx = (size_t)r.py * P.width + r.px;
            unsigned nb0 = to_u8_clamp(((double)r.n0 + 1.0) * 0.5 * 255.0);
            unsigned nb1 = to_u8_clamp(((double)r.n1 + 1.0) * 0.5 * 255.0);
            unsigned nb2 = to_u8_clamp(((double)r.n2 + 1.0) * 0.5 * 255.0);
            unsigned db = to_u8_clamp(r.depth);
            unsigned sdf16 = r.sdf & 0xffffu, it16 = r.iters & 0xffffu;
            if constexpr (kHeldEpi) {
                // keep the pixel packed in the dead ray state (q[0..2], h[0]); it is written at the warp's next refill point
                const uchar4 c1 = P.rgba ? shade_pixel<NP>(P.shader, db, nb0, nb1, nb2, sdf16, it16) : make_uchar4(0, 0, 0, 0);
                const uchar4 c2 = P.rgba2 ? shade_pixel<NP>(P.shader2, db, nb0, nb1, nb2, sdf16, it16) : make_uchar4(0, 0, 0, 0);
                r.q[0] = __uint_as_float(db | (nb0 << 8) | (nb1 << 16) | (nb2 << 24));
                r.q[1] = __uint_as_float(sdf16 | (it16 << 16));
                r.q[2] = __uint_as_float((unsigned)c1.x | ((unsigned)c1.y << 8) | ((unsigned)c1.z << 16) | ((unsigned)c1.w << 24));
                r.h[0] = __uint_as_float((unsigned)c2.x | ((unsigned)c2.y << 8) | ((unsigned)c2.z << 16) | ((unsigned)c2.w << 24));
            } else {
                P.depth[idx] = (uint8_t)db;
                P.normal[3 * idx + 0] = (uint8_t)nb0;
                P.normal[3 * idx + 1] = (uint8_t)nb1;
                P.normal[3 * idx + 2] = (uint8_t)nb2;
                P.sdf[idx] = (uint16_t)sdf16;
                P.iters[idx] = (uint16_t)it16;
                if (P.rgba) reinterpret_cast<uchar4*>(P.rgba)[idx] = shade_pixel<NP>(P.shader, db, nb0, nb1, nb2, sdf16, it16);
                if (P.rgba2) reinterpret_cast<uchar4*>(P.rgba2)[idx] = shade_pixel<NP>(P.shader2, db, nb0, nb1, nb2, sdf16, it16);
                if (P.band_flags) atomicAdd(&shBandFin[warpId][r.py / P.band_rows], 1u);  // early download, flushed per tile
            }
            if (P.depth_f32) P.depth_f32[idx] = (float)r.depth;
            if (P.depth_f64) P.depth_f64[idx] = r.depth;
            if (P.sdf_u32) P.sdf_u32[idx] = r.sdf;
            if constexpr (!NP::kExact && PK == PK_TSPHERE) {  // every evaluation is a sphere: the per-type counters are dead weight
                r.nSphere = r.sdf;
                r.nBox = 0u;
            }
            if constexpr (kSmemStats) {
                // warp-aggregated: the lanes retiring together reduce their values with REDUX and one of them updates the warp's
                // accumulators (atomics: another divergent group of the same warp may be doing the same)
                const unsigned am = __activemask();
                const bool lead = lane == (__ffs(am) - 1);
                const unsigned hitv = (r.depth < MAX_DIST) ? 1u : 0u;
                unsigned v5 = r.nBox, v6, v8 = 0u;
                if constexpr (NP::kExact) {
                    v6 = r.ex.nTorus;
                    v8 = r.ex.opFlops;
                } else {
                    v6 = r.sdf - r.nSphere - r.nBox;
                }
                const unsigned a0 = __reduce_add_sync(am, sdf16), a1 = __reduce_add_sync(am, it16), a3 = __reduce_add_sync(am, r.iters);
                const unsigned a7 = __reduce_add_sync(am, hitv);
                // counters that can exceed 2^27 per pixel (n_prims x steps) are summed in two 16-bit halves
                const unsigned long long a2 = (unsigned long long)__reduce_add_sync(am, r.sdf & 0xffffu) + ((unsigned long long)__reduce_add_sync(am, r.sdf >> 16) << 16);
                const unsigned long long a4 = (unsigned long long)__reduce_add_sync(am, r.nSphere & 0xffffu) + ((unsigned long long)__reduce_add_sync(am, r.nSphere >> 16) << 16);
                const unsigned long long a5 = (unsigned long long)__reduce_add_sync(am, v5 & 0xffffu) + ((unsigned long long)__reduce_add_sync(am, v5 >> 16) << 16);
                const unsigned long long a6 = (unsigned long long)__reduce_add_sync(am, v6 & 0xffffu) + ((unsigned long long)__reduce_add_sync(am, v6 >> 16) << 16);
                unsigned long long a8 = 0ull;
                if constexpr (NP::kExact) a8 = (unsigned long long)__reduce_add_sync(am, v8 & 0xffffu) + ((unsigned long long)__reduce_add_sync(am, v8 >> 16) << 16);
                const unsigned m0 = __reduce_max_sync(am, sdf16), m1 = __reduce_min_sync(am, sdf16), m2 = __reduce_max_sync(am, it16), m3 = __reduce_min_sync(am, it16);
                if (lead) {
                    unsigned long long* ss = shStatSum[threadIdx.x >> 5];
                    unsigned* mm = shStatMM[threadIdx.x >> 5];
                    atomicAdd(&ss[0], (unsigned long long)a0);
                    atomicAdd(&ss[1], (unsigned long long)a1);
                    atomicAdd(&ss[2], a2);
                    atomicAdd(&ss[3], (unsigned long long)a3);
                    atomicAdd(&ss[4], a4);
                    if (a5) atomicAdd(&ss[5], a5);
                    if (a6) atomicAdd(&ss[6], a6);
                    if (a7) atomicAdd(&ss[7], (unsigned long long)a7);
                    if (a8) atomicAdd(&ss[8], a8);
                    atomicMax(&mm[0], m0);
                    atomicMin(&mm[1], m1);
                    atomicMax(&mm[2], m2);
                    atomicMin(&mm[3], m3);
                }
            } else {
            st.sum_sdf += sdf16;
            st.sum_iters += it16;
            st.sum_sdf_full += r.sdf;
            st.sum_iters_full += r.iters;
            st.ev_sphere += r.nSphere;
            st.ev_box += r.nBox;
            if constexpr (NP::kExact) {
                st.ev_torus += r.ex.nTorus;
                st.op_flops += r.ex.opFlops;
            } else {
                st.ev_torus += r.sdf - r.nSphere - r.nBox;
            }
            st.n_hit += (r.depth < MAX_DIST) ? 1u : 0u;
            st.max_sdf = max(st.max_sdf, sdf16);
            st.min_sdf = min(st.min_sdf, sdf16);
            st.max_iters = max(st.max_iters, it16);
            st.min_iters = min(st.min_iters, it16);
            }
            r.phase = kHeldEpi ? PH_HELD : PH_IDLE;
        }

        // ---- (g) cooperative mode: publish whether this warp can still make progress on its own ----
        if (useQueue) {
            const unsigned pendM = __ballot_sync(kFull, r.pending);
            const unsigned actM = __ballot_sync(kFull, r.phase != PH_IDLE && !r.pending);
            const unsigned idleM = __ballot_sync(kFull, r.phase == PH_IDLE);
            const bool stuckNow = (actM == 0u) && (queueEmpty || idleM == 0u);
            const bool finishedNow = stuckNow && pendM == 0u;
            if (lane == 0) {
                if (finishedNow != wasFinished) atomicAdd(&shFinished, finishedNow ? 1u : 0xffffffffu);
                if (stuckNow != wasStuck) {
                    const unsigned prev = atomicAdd(&shStuck, stuckNow ? 1u : 0xffffffffu);
                    // last warp to get stuck: nobody can progress any more -> serve whatever is queued.  End of the frame (this
                    // warp's tile queue is empty): the SM is no longer throughput-bound, the rays still in flight are a chain of
                    // dependent passes — serve the queue as soon as P.tail_trigger warps are parked instead of all of them.
                    const unsigned needStuck = queueEmpty ? min((unsigned)kWarpsPerCta, (unsigned)P.tail_trigger) : (unsigned)kWarpsPerCta;
                    if (stuckNow && prev + 1u >= needStuck) {
                        __threadfence_block();
                        if (*(volatile unsigned*)&shTail != *(volatile unsigned*)&shHead) *(volatile unsigned*)&shGo = 1u;
                    }
                }
            }
            wasStuck = stuckNow;
            wasFinished = finishedNow;
        }
    }

#ifdef RM_PHASE_TIMING
    if (lane == 0) {
        atomicAdd(&P.stats->t_total, (unsigned long long)(clock64() - tKernel0));
        atomicAdd(&P.stats->t_search, (unsigned long long)tSearch);
        atomicAdd(&P.stats->t_barrier, (unsigned long long)tBarrier);
        atomicAdd(&P.stats->t_stuck, (unsigned long long)tStuck);
    }
#endif
    if constexpr (kTC) {
        if (useTC) {  // every warp of the CTA leaves the loop together (shFinished == all): release the TMEM columns
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (warpId == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tc.tmem), "r"(512u) : "memory");
        }
    }
    if constexpr (kHeldEpi) {  // (every path out of the loop has passed a refill point with no active lane; kept as a safety net)
        const unsigned heldM = __ballot_sync(kFull, r.phase == PH_HELD);
        if (heldM) flush_held(heldM);
    }
    if (P.tile_cost && tile >= 0 && lane == 0) P.tile_cost[tile] = (unsigned)clock() - tileT0;  // the warp's last tile
    if (P.anatomy && lane == 0) {  // frame anatomy (rm_stats: drain / tail), RM_ANATOMY=1: three same-address atomics per warp
        const unsigned long long now = globaltimer_ns();
        atomicMin(&P.stats->t_exit_min, now);
        atomicMax(&P.stats->t_exit_max, now);
    }
    publish_bands();  // pixels finalised since the warp's last tile fetch
    // ---- epilogue: diagnostics (main.ts:527-548) — warp reduce, one atomic set per warp ----
    unsigned long long s0, s1, s2, s3, s4, s5, s6, s7, s8 = 0;
    unsigned mx0, mn0, mx1, mn1;
    if constexpr (kSmemStats) {
        __syncwarp();
        const unsigned long long* ss = shStatSum[threadIdx.x >> 5];
        const unsigned* mm = shStatMM[threadIdx.x >> 5];
        s0 = ss[0], s1 = ss[1], s2 = ss[2], s3 = ss[3], s4 = ss[4], s5 = ss[5], s6 = ss[6], s7 = ss[7], s8 = ss[8];
        mx0 = mm[0], mn0 = mm[1], mx1 = mm[2], mn1 = mm[3];
    } else {
        s0 = warp_sum_u64(st.sum_sdf), s1 = warp_sum_u64(st.sum_iters);
        s2 = warp_sum_u64(st.sum_sdf_full), s3 = warp_sum_u64(st.sum_iters_full);
        s4 = warp_sum_u64(st.ev_sphere), s5 = warp_sum_u64(st.ev_box), s6 = warp_sum_u64(st.ev_torus);
        s7 = warp_sum_u64(st.n_hit);
        if constexpr (NP::kExact) s8 = warp_sum_u64(st.op_flops);
        mx0 = __reduce_max_sync(kFull, st.max_sdf), mn0 = __reduce_min_sync(kFull, st.min_sdf);
        mx1 = __reduce_max_sync(kFull, st.max_iters), mn1 = __reduce_min_sync(kFull, st.min_iters);
    }
    if (lane == 0) {
        DevStats* g = P.stats;
        atomicAdd(&g->sum_sdf, s0);
        atomicAdd(&g->sum_iters, s1);
        atomicAdd(&g->sum_sdf_full, s2);
        atomicAdd(&g->sum_iters_full, s3);
        atomicAdd(&g->evals_sphere, s4);
        atomicAdd(&g->evals_box, s5);
        atomicAdd(&g->evals_torus, s6);
        atomicAdd(&g->n_hit, s7);
        if constexpr (NP::kExact) {
            if (s8) atomicAdd(&g->op_flops, s8);
        }
        atomicMax(&g->max_sdf, mx0);
        atomicMin(&g->min_sdf, mn0);
        atomicMax(&g->max_iters, mx1);
        atomicMin(&g->min_iters, mn1);
    }
}

// Cost-ordered tile queue: one CTA turns the per-tile cycle counts a frame recorded into the tile order of the next frame of the
// same geometry.  What makes the end of a frame long is a handful of tiles that take 1-2 ms each (chains of dependent
// all-primitives passes) and happen to start late; what lets the row bands of a frame be downloaded while the rest still renders
// is that tiles are handed out in spatial order.  Both are kept:
//   1. the most expensive quarter of the tiles (by the max-scaled cost of the last frame, 256 buckets) goes FIRST, most expensive
//      first, wherever it lies in the image: every long tile starts in the first quarter of the frame;
//   2. the other tiles follow in coarse spatial order — nRuns consecutive runs of the row-major tile index, sized by the host at
//      >= 4 tiles per resident warp — most expensive first inside each run, so row bands still complete one after the other and
//      the frame ends on the cheapest tiles of its last run.
// (Measured on the 1/8 stripe share of cfg4, kernel ms / tail ms: no order 4.76 / 1.9; fully global order 4.0 / 0.65 but every
// band completes at the very end; spatial runs only 4.54 / 1.5.)  Ties keep an arbitrary order; the image does not depend on the
// schedule.
constexpr int kOrderBands = 16;
constexpr int kOrderKeys = 256 + kOrderBands * 256;
// nRuns == 0: no spatial constraint at all (frames that stay on the device: nothing downloads bands) — every tile by cost.
static __global__ void __launch_bounds__(1024) order_tiles_kernel(const unsigned* __restrict__ cost, unsigned* __restrict__ order, int n, int nRuns) {
    __shared__ unsigned shMax, shCut, hist[kOrderKeys], cursor[kOrderKeys];
    const int tid = threadIdx.x;
    if (tid == 0) shMax = 1u;
    for (int i = tid; i < kOrderKeys; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    unsigned m = 1u;
    for (int i = tid; i < n; i += blockDim.x) m = max(m, __ldg(cost + i));
    m = __reduce_max_sync(kFull, m);
    if ((tid & 31) == 0) atomicMax(&shMax, m);
    __syncthreads();
    const unsigned long long mx = shMax;
    auto bucket = [&](int i) { return 255u - (unsigned)(((unsigned long long)__ldg(cost + i) * 255ull) / mx); };  // 0 = most expensive
    // pass 1: global histogram of the cost buckets (kept in hist[0..255]) -> the bucket that closes the most expensive quarter
    for (int i = tid; i < n; i += blockDim.x) atomicAdd(&hist[bucket(i)], 1u);
    __syncthreads();
    if (tid == 0) {
        unsigned acc = 0u, cut = 0u;
        for (unsigned b = 0; b < 256u; ++b) {
            if (nRuns > 0 && acc + hist[b] > (unsigned)n / 4u) break;
            acc += hist[b];
            cut = b + 1u;  // buckets [0, cut) go first
        }
        shCut = cut;
    }
    __syncthreads();
    const unsigned cut = shCut;
    for (int i = tid; i < 256; i += blockDim.x)
        if ((unsigned)i >= cut) hist[i] = 0u;  // (their tiles are counted again under their run's keys)
    __syncthreads();
    const int perRun = nRuns > 0 ? (n + nRuns - 1) / nRuns : n;  // nRuns <= kOrderBands
    auto key = [&](int i) {
        const unsigned b = bucket(i);
        return b < cut ? b : 256u + (unsigned)(i / perRun) * 256u + b;
    };
    for (int i = tid; i < n; i += blockDim.x) {
        const unsigned k = key(i);
        if (k >= 256u) atomicAdd(&hist[k], 1u);
    }
    __syncthreads();
    if (tid < 32) {  // exclusive prefix over the counters: each lane scans a contiguous chunk, then the chunk totals
        constexpr int kPer = kOrderKeys / 32;
        static_assert(kOrderKeys % 32 == 0, "chunked scan");
        unsigned sum = 0u;
        for (int k = 0; k < kPer; ++k) sum += hist[tid * kPer + k];
        unsigned incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(kFull, incl, o);
            if (tid >= o) incl += v;
        }
        unsigned acc = incl - sum;
        for (int k = 0; k < kPer; ++k) {
            cursor[tid * kPer + k] = acc;
            acc += hist[tid * kPer + k];
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) order[atomicAdd(&cursor[key(i)], 1u)] = (unsigned)i;
}

}  // namespace rm
