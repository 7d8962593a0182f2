// ORACLE — TEST INFRASTRUCTURE ONLY (see jsnum.hpp header).  PARITY UNPINNED.
//
// oracle.cpp — C API over the restatement so Python tests / bench.py's cpu_baseline can drive it
// through ctypes.  The product (cpu_raymarcher_b200/) never links or loads this library.
// The worker entry it mirrors: src/workers/raymarchWorker.ts:33-92 (build scene + accel per job,
// set camera, run the chosen algorithm over rows [yStart, yEnd)).
#include <cstring>
#include <deque>
#include <thread>

#include "march.hpp"

using namespace orc;

struct orc_scene {
    Scene scene;
};

static Primitive prim_from_flat(int type, const float* m, const double* prm) {
    Primitive p;
    p.type = type == 3 ? (int)MANDELBULB : type;  // rm_prim_type 3 = mandelbulb: power, iterations, enableAnimation, animationSpeed
    for (int i = 0; i < 16; ++i) p.transform.e[i] = m[i];
    if (type == SPHERE) p.radius = prm[0];
    else if (type == BOX) p.halfSize = glm::v3_from(prm[0], prm[1], prm[2]);
    else if (type == 3) {
        p.power = prm[0];
        p.iterations = (int)prm[1];
        p.enableAnimation = prm[2] != 0.0;
        p.animationSpeed = prm[3];
    } else {
        p.majorRadius = prm[0];
        p.minorRadius = prm[1];
    }
    return p;
}

extern "C" {

orc_scene* orc_scene_new(void) { return new orc_scene(); }
void orc_scene_free(orc_scene* s) { delete s; }
void orc_set_length_mode(int use_hypot) { glm::length_uses_hypot() = use_hypot != 0; }

int orc_scene_load_preset(orc_scene* s, int idx) {
    std::vector<Primitive> prims;
    if (!makePreset(idx, prims)) return -1;
    s->scene.setObjects(std::move(prims));
    return 0;
}
void orc_scene_load_synthetic(orc_scene* s, int n, uint32_t seed) {
    std::vector<Primitive> prims;
    makeSyntheticSpheres(n, seed, prims);
    s->scene.setObjects(std::move(prims));
}
void orc_scene_set_prims(orc_scene* s, int n, const uint8_t* type, const float* w2l, const double* params) {
    std::vector<Primitive> prims;
    for (int i = 0; i < n; ++i) prims.push_back(prim_from_flat(type[i], w2l + 16 * i, params + 4 * i));
    s->scene.setObjects(std::move(prims));
}
int orc_scene_n_prims(orc_scene* s) { return (int)s->scene.objects.size(); }
void orc_scene_get_prims(orc_scene* s, uint8_t* type, float* w2l, double* params) {
    for (size_t i = 0; i < s->scene.objects.size(); ++i) {
        const Primitive& p = s->scene.objects[i];
        type[i] = (uint8_t)(p.type == MANDELBULB ? 3 : p.type);
        std::memcpy(w2l + 16 * i, p.transform.e, 16 * sizeof(float));
        double* q = params + 4 * i;
        q[0] = q[1] = q[2] = q[3] = 0;
        if (p.type == MANDELBULB) {
            q[0] = p.power;
            q[1] = p.iterations;
            q[2] = p.enableAnimation ? 1.0 : 0.0;
            q[3] = p.animationSpeed;
        } else if (p.type == SPHERE) q[0] = p.radius;
        else if (p.type == BOX) {
            q[0] = p.halfSize[0];
            q[1] = p.halfSize[1];
            q[2] = p.halfSize[2];
        } else {
            q[0] = p.majorRadius;
            q[1] = p.minorRadius;
        }
    }
}
// ---- operator trees as flat node arrays (same layout as rm_op_node in include/rm.h; restated here so the
// oracle stays independent of the product headers)
struct orc_op_node {
    int32_t kind;      // 0 = primitive leaf, 1 round, 2 twist, 3 smooth union, 4 smooth subtraction, 5 repetition, 6 animated translate
    int32_t child[2];  // node indices, -1 = none
    int32_t prim;      // leaf: index into the primitive arrays
    double p[4];       // round: radius | twist: amount | smooth: k | repetition: spacing xyz | animated: amplitude, speed
    float dir[4];      // animated: normalised direction
    float transform[16];
};
static_assert(sizeof(orc_op_node) == 128, "orc_op_node layout");

static Primitive prim_from_node(const orc_op_node* nodes, int i, const uint8_t* type, const float* w2l, const double* params) {
    const orc_op_node& nd = nodes[i];
    if (nd.kind == 0) return prim_from_flat(type[nd.prim], w2l + 16 * nd.prim, params + 4 * nd.prim);
    Primitive p;
    static const int kinds[7] = {-1, ROUND, TWIST, SMOOTH_UNION, SMOOTH_SUBTRACTION, REPETITION, ANIMATED_TRANSLATE};
    p.type = kinds[nd.kind];
    for (int k = 0; k < 16; ++k) p.transform.e[k] = nd.transform[k];
    p.a = std::make_shared<Primitive>(prim_from_node(nodes, nd.child[0], type, w2l, params));
    if (nd.child[1] >= 0) p.b = std::make_shared<Primitive>(prim_from_node(nodes, nd.child[1], type, w2l, params));
    switch (p.type) {
        case ROUND: p.radius = nd.p[0]; break;
        case TWIST: p.twistAmount = nd.p[0]; break;
        case SMOOTH_UNION: case SMOOTH_SUBTRACTION: p.smoothness = nd.p[0]; break;
        case REPETITION: p.spacing = glm::v3_from(nd.p[0], nd.p[1], nd.p[2]); break;
        default:
            p.amplitude = nd.p[0];
            p.speed = nd.p[1];
            for (int k = 0; k < 3; ++k) p.direction.e[k] = nd.dir[k];
            break;
    }
    return p;
}
void orc_scene_set_tree(orc_scene* s, const uint8_t* type, const float* w2l, const double* params, const void* nodes,
                        int n_objects, const int32_t* roots) {
    std::vector<Primitive> objs;
    for (int i = 0; i < n_objects; ++i) objs.push_back(prim_from_node((const orc_op_node*)nodes, roots[i], type, w2l, params));
    s->scene.setObjects(std::move(objs));
}
static void tree_count(const Primitive& p, int& nNodes, int& nPrims) {
    nNodes++;
    if (p.type <= TORUS || p.type == MANDELBULB) nPrims++;
    if (p.a) tree_count(*p.a, nNodes, nPrims);
    if (p.b) tree_count(*p.b, nNodes, nPrims);
}
void orc_scene_tree_counts(orc_scene* s, int* nNodes, int* nPrims) {
    *nNodes = *nPrims = 0;
    for (auto& p : s->scene.objects) tree_count(p, *nNodes, *nPrims);
}
static void flat_prim(const Primitive& p, uint8_t* type, float* w2l, double* q) {
    *type = (uint8_t)(p.type == MANDELBULB ? 3 : p.type);
    std::memcpy(w2l, p.transform.e, 16 * sizeof(float));
    q[0] = q[1] = q[2] = q[3] = 0;
    if (p.type == MANDELBULB) {
        q[0] = p.power;
        q[1] = p.iterations;
        q[2] = p.enableAnimation ? 1.0 : 0.0;
        q[3] = p.animationSpeed;
    } else if (p.type == SPHERE) q[0] = p.radius;
    else if (p.type == BOX) {
        q[0] = p.halfSize[0];
        q[1] = p.halfSize[1];
        q[2] = p.halfSize[2];
    } else if (p.type == TORUS) {
        q[0] = p.majorRadius;
        q[1] = p.minorRadius;
    }
}
static int tree_flat(const Primitive& p, orc_op_node* nodes, uint8_t* type, float* w2l, double* params, int& nextNode, int& nextPrim) {
    const int me = nextNode++;
    orc_op_node& nd = nodes[me];
    std::memset(&nd, 0, sizeof(nd));
    nd.child[0] = nd.child[1] = nd.prim = -1;
    std::memcpy(nd.transform, p.transform.e, sizeof(nd.transform));
    if (p.type <= TORUS || p.type == MANDELBULB) {
        nd.kind = 0;
        nd.prim = nextPrim++;
        flat_prim(p, type + nd.prim, w2l + 16 * nd.prim, params + 4 * nd.prim);
        return me;
    }
    switch (p.type) {
        case ROUND: nd.kind = 1; nd.p[0] = p.radius; break;
        case TWIST: nd.kind = 2; nd.p[0] = p.twistAmount; break;
        case SMOOTH_UNION: nd.kind = 3; nd.p[0] = p.smoothness; break;
        case SMOOTH_SUBTRACTION: nd.kind = 4; nd.p[0] = p.smoothness; break;
        case REPETITION: nd.kind = 5; nd.p[0] = p.spacing[0]; nd.p[1] = p.spacing[1]; nd.p[2] = p.spacing[2]; break;
        default:
            nd.kind = 6;
            nd.p[0] = p.amplitude;
            nd.p[1] = p.speed;
            for (int k = 0; k < 3; ++k) nd.dir[k] = p.direction.e[k];
            break;
    }
    int c0 = tree_flat(*p.a, nodes, type, w2l, params, nextNode, nextPrim);
    int c1 = p.b ? tree_flat(*p.b, nodes, type, w2l, params, nextNode, nextPrim) : -1;
    nodes[me].child[0] = c0;
    nodes[me].child[1] = c1;
    return me;
}
// pre-order per object; leaf primitives numbered in encounter order
void orc_scene_get_tree(orc_scene* s, uint8_t* type, float* w2l, double* params, void* nodes, int32_t* roots) {
    int nn = 0, np = 0;
    for (size_t i = 0; i < s->scene.objects.size(); ++i)
        roots[i] = tree_flat(s->scene.objects[i], (orc_op_node*)nodes, type, w2l, params, nn, np);
}
void orc_scene_set_time(orc_scene* s, double time) { s->scene.updateTime(time); }  // raymarcher.ts:59
// Primitive.getWorldPosition / getLocalBoundingRadius / BoundingBox.fromPrimitive of scene object i
void orc_object_geometry(orc_scene* s, int i, float* worldPos, double* radius, float* bmin, float* bmax) {
    const Primitive& p = s->scene.objects[(size_t)i];
    vec3 w = p.getWorldPosition();
    std::memcpy(worldPos, w.e, sizeof(w.e));
    *radius = p.getLocalBoundingRadius();
    BoundingBox b = BoundingBox::fromPrimitive(p);
    std::memcpy(bmin, b.min.e, sizeof(b.min.e));
    std::memcpy(bmax, b.max.e, sizeof(b.max.e));
}
double orc_js_round(double x) { return js::round(x); }

void orc_scene_build_accel(orc_scene* s, int kind) { s->scene.buildAccel(kind); }
void orc_scene_set_camera(orc_scene* s, double pitch, double yaw) { s->scene.camera.setAngles(pitch, yaw); }
// raymarcher.ts:62-67 : rot3 = mat3.fromMat4(camera.getRotationMatrix()), origin = camera.getPosition()
void orc_scene_get_camera(orc_scene* s, float* rot3, float* origin) {
    mat4 r;
    s->scene.camera.getRotationMatrix(r);
    mat3 m3;
    glm::m3_from_mat4(m3, r);
    std::memcpy(rot3, m3.e, sizeof(m3.e));
    vec3 o = glm::v3_create();
    s->scene.camera.getPosition(o);
    std::memcpy(origin, o.e, sizeof(o.e));
}
// main.ts:438-441 : analytics auto-rotate (yaw += 0.015 before each frame)
void orc_scene_rotate_camera(orc_scene* s, double dpitch, double dyaw) { s->scene.camera.rotateCamera(dpitch, dyaw); }
void orc_scene_get_angles(orc_scene* s, double* out) {
    out[0] = s->scene.camera.pitch;
    out[1] = s->scene.camera.yaw;
}

// ---- flatten the pointer trees (pre-order for the BVH, FIFO/8-consecutive-children for the octree)
static void bvh_count(const BVHNode* n, int& nodes, int& leafPrims) {
    nodes++;
    if (n->isLeaf()) leafPrims += (int)n->primitives.size();
    if (n->left) bvh_count(n->left.get(), nodes, leafPrims);
    if (n->right) bvh_count(n->right.get(), nodes, leafPrims);
}
int orc_bvh_counts(orc_scene* s, int* nNodes, int* nLeafPrims) {
    if (!s->scene.bvh) return -1;
    *nNodes = 0;
    *nLeafPrims = 0;
    bvh_count(s->scene.bvh->root.get(), *nNodes, *nLeafPrims);
    return 0;
}
static int bvh_flat(const BVHNode* n, float* bounds, int32_t* links, int32_t* leafPrims, int& nextNode, int& nextPrim) {
    int me = nextNode++;
    for (int k = 0; k < 3; ++k) {
        bounds[6 * me + k] = n->bounds.min.e[k];
        bounds[6 * me + 3 + k] = n->bounds.max.e[k];
    }
    links[4 * me + 2] = nextPrim;
    links[4 * me + 3] = 0;
    if (n->isLeaf()) {
        links[4 * me + 3] = (int)n->primitives.size();
        for (auto* p : n->primitives) leafPrims[nextPrim++] = p->index;
    }
    links[4 * me + 0] = n->left ? bvh_flat(n->left.get(), bounds, links, leafPrims, nextNode, nextPrim) : -1;
    links[4 * me + 1] = n->right ? bvh_flat(n->right.get(), bounds, links, leafPrims, nextNode, nextPrim) : -1;
    return me;
}
void orc_bvh_flatten(orc_scene* s, float* bounds, int32_t* links, int32_t* leafPrims) {
    int nn = 0, np = 0;
    bvh_flat(s->scene.bvh->root.get(), bounds, links, leafPrims, nn, np);
}
int orc_octree_counts(orc_scene* s, int* nNodes, int* nLeafPrims) {
    if (!s->scene.octree) return -1;
    int nodes = 0, prims = 0;
    std::deque<const OctreeNode*> q{s->scene.octree->root.get()};
    while (!q.empty()) {
        const OctreeNode* n = q.front();
        q.pop_front();
        nodes++;
        prims += (int)n->primitives.size();
        for (auto& c : n->children) q.push_back(c.get());
    }
    *nNodes = nodes;
    *nLeafPrims = prims;
    return 0;
}
// links: 3 per node = first_child (-1 for leaf), prim_first, prim_count
void orc_octree_flatten(orc_scene* s, float* bounds, int32_t* links, uint8_t* level, uint8_t* isEmpty, double* minDist,
                        int32_t* leafPrims) {
    std::deque<const OctreeNode*> q{s->scene.octree->root.get()};
    int me = 0, nextSlot = 1, nextPrim = 0;
    while (!q.empty()) {
        const OctreeNode* n = q.front();
        q.pop_front();
        for (int k = 0; k < 3; ++k) {
            bounds[6 * me + k] = n->bounds.min.e[k];
            bounds[6 * me + 3 + k] = n->bounds.max.e[k];
        }
        level[me] = (uint8_t)n->level;
        isEmpty[me] = n->isEmpty ? 1 : 0;
        minDist[me] = n->minDistance;
        links[3 * me + 1] = nextPrim;
        links[3 * me + 2] = (int)n->primitives.size();
        for (auto* p : n->primitives) leafPrims[nextPrim++] = p->index;
        if (n->hasChildren) {
            links[3 * me + 0] = nextSlot;
            nextSlot += 8;
            for (auto& c : n->children) q.push_back(c.get());
        } else {
            links[3 * me + 0] = -1;
        }
        me++;
    }
}

// ---- render (raymarchWorker.ts:33-92 minus the scene construction, which the caller did above)
void orc_render(orc_scene* s, int algo, double stepSize, double overshoot, int width, int height, int yStart, int yEnd,
                uint8_t* depth, uint8_t* normal, uint16_t* sdf, uint16_t* iters, double* depthF64, uint32_t* sdfFull,
                uint32_t* itersFull, int nthreads) {
    MarchParams mp;
    mp.algo = algo;
    mp.stepSize = stepSize;
    mp.overshootFactor = overshoot;
    Raymarcher rm(s->scene, mp);
    if (nthreads <= 1) {
        rm.runRows(width, height, yStart, yEnd, 0, 1, depth, normal, sdf, iters, depthF64, sdfFull, itersFull);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&, t]() {
            rm.runRows(width, height, yStart, yEnd, t, nthreads, depth, normal, sdf, iters, depthF64, sdfFull, itersFull);
        });
    for (auto& t : th) t.join();
}

// Bounded sample for CPU-baseline timing: only the listed image rows, written compactly (row k of the
// output = image row rows[k]).  Threads take rows round-robin.
void orc_render_rows(orc_scene* s, int algo, double stepSize, double overshoot, int width, int height, const int* rows,
                     int nrows, uint8_t* depth, uint8_t* normal, uint16_t* sdf, uint16_t* iters, double* depthF64,
                     uint32_t* sdfFull, uint32_t* itersFull, int nthreads) {
    MarchParams mp;
    mp.algo = algo;
    mp.stepSize = stepSize;
    mp.overshootFactor = overshoot;
    Raymarcher rm(s->scene, mp);
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&, t]() {
            for (int k = t; k < nrows; k += nthreads)
                rm.runRow(width, height, rows[k], k, depth, normal, sdf, iters, depthF64, sdfFull, itersFull);
        });
    for (auto& t : th) t.join();
}

void orc_shade(int shader, uint8_t* rgba, const uint8_t* depth, const uint8_t* normal, const uint16_t* sdf,
               const uint16_t* iters, int w, int h) {
    size_t n = (size_t)w * h;
    switch (shader) {
        case SH_PHONG: shadePhong(rgba, depth, normal, n); break;
        case SH_SDF_HEAT: shadeHeat(rgba, sdf, n); break;
        case SH_ITER_HEAT: shadeHeat(rgba, iters, n); break;
        default: shadeNormal(rgba, normal, n); break;
    }
}
// out = {total SDF calls, max, min, total iterations}  (main.ts:527-543)
void orc_stats(const uint16_t* sdf, const uint16_t* iters, size_t n, double* out) {
    FrameStats st = frameStats(sdf, iters, n);
    out[0] = st.totalSDFCalls;
    out[1] = st.maxSDFCalls;
    out[2] = st.minSDFCalls;
    out[3] = st.totalIterations;
}

// ---- unit probes for tests
double orc_hypot3(double a, double b, double c) { return js::hypot3(a, b, c); }
int orc_to_u8(double x) { return js::to_u8_clamp(x); }
double orc_min2(double a, double b) { return js::min2(a, b); }
double orc_max2(double a, double b) { return js::max2(a, b); }
int orc_mat4_invert(const float* in, float* out) {
    mat4 a, o = glm::m4_create();
    std::memcpy(a.e, in, sizeof(a.e));
    bool ok = glm::m4_invert(o, a);
    std::memcpy(out, o.e, sizeof(o.e));
    return ok ? 0 : -1;
}
void orc_get_transform(double x, double y, double z, const float* rot_or_null, float* out) {
    vec3 r;
    if (rot_or_null) r = vec3{{rot_or_null[0], rot_or_null[1], rot_or_null[2]}};
    mat4 m = getTransform(x, y, z, rot_or_null ? &r : nullptr);
    std::memcpy(out, m.e, sizeof(m.e));
}
double orc_prim_sdf(orc_scene* s, int prim, const float* p) {
    vec3 v{{p[0], p[1], p[2]}};
    return s->scene.objects[prim].sdf(v);
}
double orc_scene_distance(orc_scene* s, const float* p, uint32_t* count) {
    vec3 v{{p[0], p[1], p[2]}};
    uint32_t c = 0;
    double d = s->scene.getDistance(v, c);
    *count = c;
    return d;
}
double orc_mulberry32(uint32_t seed, int nth) {
    Mulberry32 r(seed);
    double v = 0;
    for (int i = 0; i <= nth; ++i) v = r.next();
    return v;
}

}  // extern "C"
