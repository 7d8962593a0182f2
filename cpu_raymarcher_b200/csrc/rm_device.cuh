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
//     (STEP -> scene-distance query -> consume -> ... -> 4 normal queries -> finalize); lanes whose
//     ray terminated are found with __ballot_sync and refilled with fresh pixels immediately, so
//     divergent ray lengths do not idle lanes; all scene-distance queries of a warp — march steps,
//     normal taps, V3 bridging taps — funnel through ONE code site;
//   * a query is first resolved through the acceleration structure (few primitives, lane-local);
//     queries that need every primitive (no acceleration structure, the BVH "empty candidate set"
//     fallback of scene.ts:173, points outside the octree root) run the dense all-primitives loop
//     in which every lane reads the same primitive record (one broadcast load per warp per primitive);
//   * per-pixel outputs are quantised exactly like the reference's typed arrays and the diagnostics
//     are reduced in the epilogue (warp reduce, one atomic set per warp).
//
// Instantiated twice over the FIELD model (rm_numeric.cuh): NumJS (fp64, bit-exact) and NumFast (fp32).
// Control arithmetic is double / non-fused in both.
#pragma once
#include "rm_numeric.cuh"
#include "rm_types.h"

namespace rm {

constexpr unsigned kFull = 0xffffffffu;

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
    PH_FINAL  // march + normal done: quantise and store
};

// ------------------------------------------------------------------------------------------
// Primitive SDFs
// ------------------------------------------------------------------------------------------

// Exact model: Primitive.sdf (primitive.ts:33-39) = f32(transformMat4(p, M)) then localSdf.
RM_DEV double prim_sdf_exact(const DevScene& sc, int j, double x, double y, double z, int length_sqrt) {
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
    } else {  // torus.ts:14-25
        double qx = sqrt(lx * lx + lz * lz) - prm[0];
        double qy = ly;
        return sqrt(qx * qx + qy * qy) - prm[1];
    }
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
RM_DEV float prim_sdf_fast_tsphere(const float4* __restrict__ rec, int j, float x, float y, float z) {
    const float4 s = __ldg(rec + j);
    float lx = x + s.x, ly = y + s.y, lz = z + s.z;
    return NumFast::sqrt_(fmaf(lx, lx, fmaf(ly, ly, lz * lz))) - s.w;
}

// One primitive evaluation at the f32 sample point q; nSphere/nBox count evaluations by type.
template <class NP, int PK>
RM_DEV typename NP::F prim_sdf(const RenderParams& P, int j, const float q[3], unsigned& nSphere, unsigned& nBox) {
    if constexpr (NP::kExact) {
        int type = P.scene.type[j];
        nSphere += (type == RM_PRIM_SPHERE);
        nBox += (type == RM_PRIM_BOX);
        return prim_sdf_exact(P.scene, j, (double)q[0], (double)q[1], (double)q[2], P.length_sqrt);
    } else if constexpr (PK == PK_TSPHERE) {
        nSphere += 1;
        return prim_sdf_fast_tsphere(P.scene.rec, j, q[0], q[1], q[2]);
    } else {
        int type;
        float d = prim_sdf_fast_general(P.scene.rec, j, q[0], q[1], q[2], type);
        nSphere += (type == RM_PRIM_SPHERE);
        nBox += (type == RM_PRIM_BOX);
        return d;
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

// Dense all-primitives evaluation: closest = min over every primitive (scene.ts:183-189), starting
// from MAX_DIST = 10.  Every lane walks the same primitive index -> broadcast loads.
//   exact model: the reference's Math.min chain in doubles;
//   fast model : fp32 SEARCH for the nearest primitive (the FLOP-dominant N-primitive loop), then ONE
//                fp64 evaluation of that primitive ("f32 search, f64 polish").  The hot loop keeps only
//                a running min per 32-primitive chunk (1 FMNMX per evaluation); the winning chunk is
//                re-scanned once to recover the index.
constexpr int kChunk = 32;
template <class NP, int PK>
RM_DEV double scene_all_prims(const RenderParams& P, const float q[3]) {
    const int n = P.scene.n_prims;
    if constexpr (NP::kExact) {
        double closest = 10.0;
        for (int j = 0; j < n; ++j)
            closest = jsmin(prim_sdf_exact(P.scene, j, (double)q[0], (double)q[1], (double)q[2], P.length_sqrt), closest);
        return closest;
    } else {
        const float4* __restrict__ rec = P.scene.rec;
        float best = 10.f;
        int bestChunk = -1, bestIdx = -1;
        int j = 0;
        for (; j + kChunk <= n; j += kChunk) {
            float m = prim_sdf_f32<PK>(rec, j, q);
#pragma unroll 8
            for (int k = 1; k < kChunk; ++k) m = fminf(m, prim_sdf_f32<PK>(rec, j + k, q));
            if (m < best) {
                best = m;
                bestChunk = j;
            }
        }
        for (; j < n; ++j) {  // tail (and the whole scene when n < 32)
            float d = prim_sdf_f32<PK>(rec, j, q);
            if (d < best) {
                best = d;
                bestIdx = j;
                bestChunk = -1;
            }
        }
        if (bestChunk >= 0) {
            for (int k = 0; k < kChunk; ++k)
                if (prim_sdf_f32<PK>(rec, bestChunk + k, q) == best) {
                    bestIdx = bestChunk + k;
                    break;
                }
        }
        if (bestIdx < 0) return 10.0;  // nothing closer than MAX_DIST
        return jsmin(prim_sdf_exact(P.scene, bestIdx, (double)q[0], (double)q[1], (double)q[2], 1), 10.0);
    }
}

// Candidate-list evaluation (octree leaf / BVH leaf): running (min, argmin) in the field model.
template <class NP, int PK>
RM_DEV void leaf_prims(const RenderParams& P, const int32_t* __restrict__ lp, int pc, const float q[3], double& distExact,
                       float& distF32, int& argmin, unsigned& nSphere, unsigned& nBox) {
    for (int k = 0; k < pc; ++k) {
        const int j = lp[k];
        if constexpr (NP::kExact) {
            distExact = jsmin(prim_sdf<NP, PK>(P, j, q, nSphere, nBox), distExact);
        } else {
            float d = (float)prim_sdf<NP, PK>(P, j, q, nSphere, nBox);
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
    unsigned nSphere, nBox;    // evaluations by type (torus = sdf - nSphere - nBox)
    int i;                     // march loop index
    int phase;
    int px, py;  // pixel (x, band-local y)
    bool done;   // rayMarch has returned (depth is valid)
    bool pending;  // parked on the dense all-primitives pass
    int cur, nIv;  // BVH interval cursor (bvh.ts:204-240)
};

// BVH per-ray interval list.  The cursor advances at most one slot per march-loop index
// (bvh.ts:222-236) and there are at most MAX_STEPS indices, so only the first MAX_STEPS+1 entries of
// the stably sorted list can ever be read; exactly those are kept, by bounded stable insertion.
struct IvList {
    double enter[kMaxStepsFixed + 1];
    double exit_[kMaxStepsFixed + 1];
};

// BVH.findRayIntersections + onRayMarchStart (bvh.ts:126-202).  Returns the number of kept intervals.
static __device__ __noinline__ int bvh_collect(const rm_bvh_node* __restrict__ nodes, const double o[3], const float d[3], IvList& iv,
                                        int cap) {
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

// BVH.onRayMarchStep (bvh.ts:204-240).
template <class NP>
RM_DEV double bvh_step(Ray<NP>& r, const IvList& iv) {
    if (r.cur >= r.nIv) return -1.0;
    double e = iv.enter[r.cur];
    if (r.t < e) return e - r.t;
    if (r.t > iv.exit_[r.cur]) {
        r.cur++;
        if (r.cur < r.nIv) {
            double ne = iv.enter[r.cur];
            if (ne > r.t) return ne - r.t;
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
    unsigned long long ev_sphere = 0, ev_box = 0, ev_torus = 0, n_hit = 0;
    unsigned max_sdf = 0, min_sdf = 0xffffffffu, max_iters = 0, min_iters = 0xffffffffu;
};

RM_DEV unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

template <class NP, int ACCEL, int PK>
__global__ void __launch_bounds__(128) render_kernel(const __grid_constant__ RenderParams P) {
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;

    const double o[3] = {(double)P.origin[0], (double)P.origin[1], (double)P.origin[2]};
    const double MAX_DIST = 10.0, EPSILON = 0.001;
    const int alg = P.algorithm;
    const bool hitOnly = (alg == RM_ALG_FIXED_STEP || alg == RM_ALG_ADAPTIVE_STEP);  // return hit ? t : MAX_DIST
    const int maxSteps = hitOnly ? kMaxStepsFixed : kMaxStepsSphere;
    const double stepSize = P.step_size, overshoot = P.overshoot;
    const int bandH = P.y_end - P.y_start;

    Ray<NP> r;
    r.phase = PH_IDLE;
    r.cur = 0;
    r.nIv = 0;
    IvList iv;  // only touched when ACCEL == BVH (lives in local memory)
    LaneStats st;

    // warp-uniform work-queue cursor
    int tile = -1, tilePos = kTileW * kTileH;
    bool queueEmpty = false;

    // Warp scheduler thresholds.  Expensive, warp-serialising stages are deferred until enough lanes
    // want them: the all-primitives pass (cost ~ n_prims per execution, independent of how many lanes
    // take part) runs when kBfLanes lanes are parked on it, BVH ray set-up (a full tree traversal) when
    // kInitLanes lanes are free.  Both are forced as soon as no lane can make progress otherwise.
    const int bfLanes = (ACCEL == RM_ACCEL_NONE) ? 1 : (P.scene.n_prims >= 2048 ? 24 : (P.scene.n_prims >= 256 ? 12 : 1));
    const int initLanes = (ACCEL == RM_ACCEL_BVH) ? (P.scene.n_prims >= 256 ? 16 : 4) : 1;
    r.pending = false;

    for (;;) {
        // ---- (a) refill: lanes without a ray claim the next pixels of the warp's current tile ----
        unsigned idle = __ballot_sync(kFull, r.phase == PH_IDLE);
        {
            const int nIdle = __popc(idle);
            const int nPend = __popc(__ballot_sync(kFull, r.pending));
            const int nAct = 32 - nIdle - nPend;
            const bool doRefill = (nAct == 0) ? (nPend < bfLanes) : (nIdle >= initLanes);
            if (!doRefill) idle = 0u;
        }
        while (idle && !queueEmpty) {
            if (tilePos >= kTileW * kTileH) {
                unsigned t = 0;
                if (lane == 0) t = atomicAdd(&P.stats->queue, 1u);
                t = __shfl_sync(kFull, t, 0);
                if ((int)t >= P.n_tiles) {
                    queueEmpty = true;
                    break;
                }
                tile = (int)t;
                tilePos = 0;
            }
            int avail = kTileW * kTileH - tilePos;
            int nIdle = __popc(idle);
            int take = avail < nIdle ? avail : nIdle;
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
        if (queueEmpty && __ballot_sync(kFull, r.phase != PH_IDLE) == 0u) break;  // queue drained and every ray retired

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
            r.i = 0;
            r.done = false;
            r.cur = 0;
            r.depth = 0.0;
            r.phase = PH_STEP;
            if constexpr (ACCEL == RM_ACCEL_BVH) {
                r.nIv = bvh_collect(P.scene.bvh, o, r.d, iv, maxSteps + 1);
                if (r.nIv == 0) {  // {terminate:true} -> return MAX_DIST (sphereTracer.ts:38-40)
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
                    if constexpr (ACCEL == RM_ACCEL_BVH) skip = bvh_step<NP>(r, iv);
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
        double dd = 10.0;      // the query result handed to the control logic
        float distF = 10.f;    // fast model: running fp32 min over the candidate primitives
        int argmin = -1;       // fast model: its primitive
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
                        leaf_prims<NP, PK>(P, P.scene.leaf_prims + nd->prim_first, pc, r.q, dd, distF, argmin, r.nSphere, r.nBox);
                        cnt = (unsigned)pc;
                        polish = true;
                    } else if (nd->is_empty) {
                        dd = jsmin(10.0, nd->min_distance * 0.99);
                    }
                }
            } else {  // BVH.getPrimitivesAt (bvh.ts:95-121): every leaf whose box contains p, left before right
                const rm_bvh_node* nodes = P.scene.bvh;
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
                        leaf_prims<NP, PK>(P, P.scene.leaf_prims + nd->prim_first, pc, r.q, dd, distF, argmin, r.nSphere, r.nBox);
                        cnt += (unsigned)pc;
                    } else {
                        if (right >= 0 && sp < kBvhStack) stack[sp++] = right;  // popped after left
                        if (left >= 0 && sp < kBvhStack) stack[sp++] = left;
                    }
                }
                if (cnt == 0) needAll = true;  // candidates.length === 0 -> every primitive (scene.ts:173)
                else polish = true;
            }
        }
        if constexpr (!NP::kExact) {
            // fast model: one fp64 evaluation of the nearest candidate found by the fp32 search
            if (polish && argmin >= 0)
                dd = jsmin(prim_sdf_exact(P.scene, argmin, (double)r.q[0], (double)r.q[1], (double)r.q[2], 1), 10.0);
        }
        // dense all-primitives pass: lanes that need it park (r.pending) until enough of them have
        // gathered or nothing else in the warp can make progress
        r.pending = r.pending || needAll;
        {
            const unsigned pend = __ballot_sync(kFull, r.pending);
            const unsigned others = __ballot_sync(kFull, r.phase != PH_IDLE && !r.pending);
            const bool doAll = pend != 0u && (__popc(pend) >= bfLanes || others == 0u);
            if (doAll && r.pending) {
                dd = scene_all_prims<NP, PK>(P, r.q);
                cnt = (unsigned)P.scene.n_prims;
                r.nSphere += P.scene.type_hist[0];
                r.nBox += P.scene.type_hist[1];
                r.pending = false;
            }
        }

        // ---- (e) consume the query result ----
        if (waiting && !r.pending) {
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
            P.depth[idx] = (uint8_t)db;
            P.normal[3 * idx + 0] = (uint8_t)nb0;
            P.normal[3 * idx + 1] = (uint8_t)nb1;
            P.normal[3 * idx + 2] = (uint8_t)nb2;
            P.sdf[idx] = (uint16_t)sdf16;
            P.iters[idx] = (uint16_t)it16;
            if (P.rgba) reinterpret_cast<uchar4*>(P.rgba)[idx] = shade_pixel<NP>(P.shader, db, nb0, nb1, nb2, sdf16, it16);
            if (P.rgba2) reinterpret_cast<uchar4*>(P.rgba2)[idx] = shade_pixel<NP>(P.shader2, db, nb0, nb1, nb2, sdf16, it16);
            if (P.depth_f32) P.depth_f32[idx] = (float)r.depth;
            if (P.depth_f64) P.depth_f64[idx] = r.depth;
            if (P.sdf_u32) P.sdf_u32[idx] = r.sdf;
            st.sum_sdf += sdf16;
            st.sum_iters += it16;
            st.sum_sdf_full += r.sdf;
            st.sum_iters_full += r.iters;
            st.ev_sphere += r.nSphere;
            st.ev_box += r.nBox;
            st.ev_torus += r.sdf - r.nSphere - r.nBox;
            st.n_hit += (r.depth < MAX_DIST) ? 1u : 0u;
            st.max_sdf = max(st.max_sdf, sdf16);
            st.min_sdf = min(st.min_sdf, sdf16);
            st.max_iters = max(st.max_iters, it16);
            st.min_iters = min(st.min_iters, it16);
            r.phase = PH_IDLE;
        }
    }

    // ---- epilogue: diagnostics (main.ts:527-548) — warp reduce, one atomic set per warp ----
    unsigned long long s0 = warp_sum_u64(st.sum_sdf), s1 = warp_sum_u64(st.sum_iters);
    unsigned long long s2 = warp_sum_u64(st.sum_sdf_full), s3 = warp_sum_u64(st.sum_iters_full);
    unsigned long long s4 = warp_sum_u64(st.ev_sphere), s5 = warp_sum_u64(st.ev_box), s6 = warp_sum_u64(st.ev_torus);
    unsigned long long s7 = warp_sum_u64(st.n_hit);
    unsigned mx0 = __reduce_max_sync(kFull, st.max_sdf), mn0 = __reduce_min_sync(kFull, st.min_sdf);
    unsigned mx1 = __reduce_max_sync(kFull, st.max_iters), mn1 = __reduce_min_sync(kFull, st.min_iters);
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
        atomicMax(&g->max_sdf, mx0);
        atomicMin(&g->min_sdf, mn0);
        atomicMax(&g->max_iters, mx1);
        atomicMin(&g->min_iters, mn1);
    }
}

}  // namespace rm
