// rm_build.cpp — native host builders for the flattened BVH / octree (compiled -ffp-contract=off).
//
// Result-identical to the reference's builders but with sane complexity: the reference re-inverts
// every primitive's mat4 inside its sort comparator and per tree level (bvh.ts:66-70,84-85;
// boundingBox.ts:133-154) and rebuilds per job per frame (raymarchWorker.ts:37-38).  Here the world
// position and padded box of every primitive are computed once; the rest is index shuffling.
//   BVH    : src/acceleration_structures/bvh.ts:29-92     (median split, stable sort, depth<=20, leaf<=2)
//   Octree : src/acceleration_structures/octree.ts:36-118 (fixed +-10 root from scene.ts:81-85, z-y-x
//            children, closed-interval assignment, depth<=6, leaf<=4) and :149-191 (minDistance)
// Arithmetic model: JS doubles, f32 stores.  This is product code and does
// not share sources with oracle/.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <limits>
#include <system_error>
#include <thread>
#include <vector>

#include "rm_host.h"

namespace rm {

static inline float f32(double x) { return (float)x; }

// V8 Math.hypot (three arguments).
static double hypot3(double a, double b, double c) {
    double v[3] = {a, b, c}, av[3] = {0, 0, 0};
    bool nan = false;
    double mx = 0;
    for (int i = 0; i < 3; ++i) {
        if (std::isnan(v[i])) nan = true;
        else {
            av[i] = std::fabs(v[i]);
            if (av[i] > mx) mx = av[i];
        }
    }
    if (mx == std::numeric_limits<double>::infinity()) return mx;
    if (nan) return std::numeric_limits<double>::quiet_NaN();
    if (mx == 0) return 0;
    double sum = 0, comp = 0;
    for (int i = 0; i < 3; ++i) {
        double n = av[i] / mx;
        double summand = (n * n) - comp;
        double prelim = sum + summand;
        comp = (prelim - sum) - summand;
        sum = prelim;
    }
    return std::sqrt(sum) * mx;
}

// gl-matrix mat4.invert on f32 storage (double arithmetic, f32 stores).  false when det == 0.
static bool mat4_invert(const float* a, float* out) {
    double a00 = a[0], a01 = a[1], a02 = a[2], a03 = a[3], a10 = a[4], a11 = a[5], a12 = a[6], a13 = a[7];
    double a20 = a[8], a21 = a[9], a22 = a[10], a23 = a[11], a30 = a[12], a31 = a[13], a32 = a[14], a33 = a[15];
    double b00 = a00 * a11 - a01 * a10, b01 = a00 * a12 - a02 * a10, b02 = a00 * a13 - a03 * a10;
    double b03 = a01 * a12 - a02 * a11, b04 = a01 * a13 - a03 * a11, b05 = a02 * a13 - a03 * a12;
    double b06 = a20 * a31 - a21 * a30, b07 = a20 * a32 - a22 * a30, b08 = a20 * a33 - a23 * a30;
    double b09 = a21 * a32 - a22 * a31, b10 = a21 * a33 - a23 * a31, b11 = a22 * a33 - a23 * a32;
    double det = b00 * b11 - b01 * b10 + b02 * b09 + b03 * b08 - b04 * b07 + b05 * b06;
    if (det == 0.0 || std::isnan(det)) return false;
    det = 1.0 / det;
    out[0] = f32((a11 * b11 - a12 * b10 + a13 * b09) * det);
    out[1] = f32((a02 * b10 - a01 * b11 - a03 * b09) * det);
    out[2] = f32((a31 * b05 - a32 * b04 + a33 * b03) * det);
    out[3] = f32((a22 * b04 - a21 * b05 - a23 * b03) * det);
    out[4] = f32((a12 * b08 - a10 * b11 - a13 * b07) * det);
    out[5] = f32((a00 * b11 - a02 * b08 + a03 * b07) * det);
    out[6] = f32((a32 * b02 - a30 * b05 - a33 * b01) * det);
    out[7] = f32((a20 * b05 - a22 * b02 + a23 * b01) * det);
    out[8] = f32((a10 * b10 - a11 * b08 + a13 * b06) * det);
    out[9] = f32((a01 * b08 - a00 * b10 - a03 * b06) * det);
    out[10] = f32((a30 * b04 - a31 * b02 + a33 * b00) * det);
    out[11] = f32((a21 * b02 - a20 * b04 - a23 * b00) * det);
    out[12] = f32((a11 * b07 - a10 * b09 - a12 * b06) * det);
    out[13] = f32((a00 * b09 - a01 * b07 + a02 * b06) * det);
    out[14] = f32((a31 * b01 - a30 * b03 - a32 * b00) * det);
    out[15] = f32((a20 * b03 - a21 * b01 + a22 * b00) * det);
    return true;
}

void compute_prim_geometry(int32_t n, const uint8_t* type, const float* w2l, const double* params, unsigned flags,
                           std::vector<PrimGeom>& out) {
    out.resize((size_t)n);
    const bool lengthSqrt = (flags & RM_F_LENGTH_SQRT) != 0;
    for (int32_t i = 0; i < n; ++i) {
        const float* m = w2l + 16 * (size_t)i;
        float inv[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};  // mat4.create(); untouched if singular
        bool ok = mat4_invert(m, inv);
        PrimGeom& g = out[(size_t)i];
        // Primitive.getWorldPosition (primitive.ts:20-30)
        g.world[0] = inv[12];
        g.world[1] = inv[13];
        g.world[2] = inv[14];
        // getLocalBoundingRadius (sphere.ts:16, box.ts:32-34, torus.ts:27-29)
        const double* q = params + 4 * (size_t)i;
        double localRadius;
        if (type[i] == RM_PRIM_SPHERE) localRadius = q[0];
        else if (type[i] == RM_PRIM_BOX) {
            double hx = (double)f32(q[0]), hy = (double)f32(q[1]), hz = (double)f32(q[2]);
            localRadius = lengthSqrt ? std::sqrt(hx * hx + hy * hy + hz * hz) : hypot3(hx, hy, hz);
        } else if (type[i] == RM_PRIM_MANDELBULB) localRadius = 2.5;  // mandelbulb.ts:80-83
        else localRadius = q[0] + q[1];
        // BoundingBox.fromPrimitive (boundingBox.ts:133-154)
        const float* s = ok ? inv : m;
        double scaleX = hypot3(s[0], s[1], s[2]);
        double scaleY = hypot3(s[4], s[5], s[6]);
        double scaleZ = hypot3(s[8], s[9], s[10]);
        double maxScale = std::max(std::max(scaleX, scaleY), scaleZ);
        if (std::isnan(scaleX) || std::isnan(scaleY) || std::isnan(scaleZ)) maxScale = std::numeric_limits<double>::quiet_NaN();
        double r = localRadius * maxScale * 1.5;
        for (int k = 0; k < 3; ++k) {
            g.bmin[k] = f32((double)g.world[k] - r);
            g.bmax[k] = f32((double)g.world[k] + r);
        }
    }
}

// ------------------------------------------------------------------------------------------ BVH
namespace {
// Median-split BVH, built in place.  Every split of bvh.ts is by COUNT (mid = floor(n / 2)) and every object ends up in exactly
// one leaf, so the shape of the tree — hence every node's index in the reference's pre-order numbering and every leaf's range in
// the leaf-index array — follows from n and the depth alone: a subtree over n objects at depth d has count_nodes(n, d) nodes and
// its leaves cover a contiguous run of n entries.  That makes the two children independent jobs (the top levels are handed to
// threads) and lets all nodes share one permutation array, sorted a segment at a time on (key, index) pairs instead of through a
// comparator that chases indices into the geometry.
struct BvhBuilder {
    const PrimGeom* geom;
    rm_bvh_node* nodes;
    int32_t* perm;  // becomes leafPrims
    static constexpr int kMaxDepth = 20, kMaxPrimsPerNode = 2;  // bvh.ts:32-33
    static constexpr size_t kParallelMin = 8192;                // segments at least this long fork while depth < kForkDepth
    static constexpr int kForkDepth = 4;

    static bool is_leaf(size_t n, int depth) { return depth >= kMaxDepth || n <= (size_t)kMaxPrimsPerNode; }
    // a subtree's node count depends on (n, depth) only; per depth at most two distinct n occur (floor / ceil halves)
    struct CountMemo {
        size_t n[kMaxDepth + 1][2];
        int32_t c[kMaxDepth + 1][2];
        int used[kMaxDepth + 1];
        CountMemo() { std::memset(used, 0, sizeof(used)); }
    };
    static int32_t count_nodes(size_t n, int depth, CountMemo& m) {
        if (is_leaf(n, depth)) return 1;
        for (int i = 0; i < m.used[depth]; ++i)
            if (m.n[depth][i] == n) return m.c[depth][i];
        const int32_t c = 1 + count_nodes(n / 2, depth + 1, m) + count_nodes(n - n / 2, depth + 1, m);
        if (m.used[depth] < 2) {
            m.n[depth][m.used[depth]] = n;
            m.c[depth][m.used[depth]++] = c;
        }
        return c;
    }
    void bounds_of(size_t lo, size_t hi, float* bmin, float* bmax) const {  // computeBounds, folded in array order
        for (int k = 0; k < 3; ++k) {
            bmin[k] = geom[(size_t)perm[lo]].bmin[k];
            bmax[k] = geom[(size_t)perm[lo]].bmax[k];
        }
        for (size_t i = lo + 1; i < hi; ++i) {
            const PrimGeom& g = geom[(size_t)perm[i]];
            for (int k = 0; k < 3; ++k) {
                bmin[k] = std::min(bmin[k], g.bmin[k]);
                bmax[k] = std::max(bmax[k], g.bmax[k]);
            }
        }
    }
    void build(size_t lo, size_t hi, const float* bmin, const float* bmax, int depth, int32_t me, CountMemo& memo,
               std::vector<std::pair<float, int32_t>>& keys) {
        rm_bvh_node& nd = nodes[(size_t)me];
        for (int k = 0; k < 3; ++k) {
            nd.bmin[k] = bmin[k];
            nd.bmax[k] = bmax[k];
        }
        nd.left = nd.right = -1;
        nd.prim_first = (int32_t)lo;  // = leafPrims.length when the reference creates this node
        nd.prim_count = 0;
        const size_t n = hi - lo;
        if (is_leaf(n, depth)) {
            nd.prim_count = (int32_t)n;
            return;
        }
        // longest axis of the f32 size vector (bvh.ts:57-63)
        float size[3];
        for (int k = 0; k < 3; ++k) size[k] = f32((double)bmax[k] - (double)bmin[k]);
        int axis = 0;
        if (size[1] > size[0]) axis = 1;
        if (size[2] > size[axis]) axis = 2;
        // Array.prototype.sort is stable (bvh.ts:66-70)
        keys.resize(n);
        for (size_t i = 0; i < n; ++i) keys[i] = {geom[(size_t)perm[lo + i]].world[axis], perm[lo + i]};
        std::stable_sort(keys.begin(), keys.end(), [](const std::pair<float, int32_t>& a, const std::pair<float, int32_t>& b) { return a.first < b.first; });
        for (size_t i = 0; i < n; ++i) perm[lo + i] = keys[i].second;
        const size_t mid = lo + n / 2;
        float lmin[3], lmax[3], rmin[3], rmax[3];
        bounds_of(lo, mid, lmin, lmax);
        bounds_of(mid, hi, rmin, rmax);
        const int32_t l = me + 1, r = l + count_nodes(mid - lo, depth + 1, memo);
        nd.left = l;
        nd.right = r;
        std::thread t;
        if (depth < kForkDepth && n >= kParallelMin) {
            try {
                t = std::thread([&, l] {
                    CountMemo m2;
                    std::vector<std::pair<float, int32_t>> k2;
                    build(lo, mid, lmin, lmax, depth + 1, l, m2, k2);
                });
            } catch (const std::system_error&) {  // no thread to be had: build the left subtree here as well
            }
        }
        if (t.joinable()) {
            build(mid, hi, rmin, rmax, depth + 1, r, memo, keys);
            t.join();
        } else {
            build(lo, mid, lmin, lmax, depth + 1, l, memo, keys);
            build(mid, hi, rmin, rmax, depth + 1, r, memo, keys);
        }
    }
};
}  // namespace

void build_bvh(const std::vector<PrimGeom>& geom, std::vector<rm_bvh_node>& nodes, std::vector<int32_t>& leafPrims) {
    const size_t n = geom.size();
    leafPrims.resize(n);
    for (size_t i = 0; i < n; ++i) leafPrims[i] = (int32_t)i;
    BvhBuilder::CountMemo memo;
    nodes.assign((size_t)BvhBuilder::count_nodes(n, 0, memo), rm_bvh_node{});
    BvhBuilder b{geom.data(), nodes.data(), leafPrims.data()};
    float bmin[3] = {0.f, 0.f, 0.f}, bmax[3] = {0.f, 0.f, 0.f};  // computeBounds of an empty scene
    if (n > 0) b.bounds_of(0, n, bmin, bmax);
    std::vector<std::pair<float, int32_t>> keys;
    b.build(0, n, bmin, bmax, 0, 0, memo, keys);
}

// ------------------------------------------------------------------------------------ leaf grid
void build_leaf_grid(const std::vector<rm_bvh_node>& nodes, LeafGrid& g) {
    g = LeafGrid();
    if (nodes.empty()) {
        g.cell_start.assign(2, 0u);
        return;
    }
    std::vector<int32_t> leafNodes;
    for (size_t i = 0; i < nodes.size(); ++i)
        if (nodes[i].left < 0 && nodes[i].right < 0 && nodes[i].prim_count > 0) leafNodes.push_back((int32_t)i);
    const rm_bvh_node& root = nodes[0];
    // resolution: about 1.75 * cbrt(#leaves) cells along the longest axis, at most 128 (ranges pack into bytes)
    double ext[3], maxExt = 0;
    for (int k = 0; k < 3; ++k) {
        ext[k] = (double)root.bmax[k] - (double)root.bmin[k];
        if (ext[k] > maxExt) maxExt = ext[k];
    }
    double factor = 1.75;  // measured flat optimum 1.25 .. 2.5 on config 4 (10k and 100k spheres)
    if (const char* e = std::getenv("RM_GRID_FACTOR")) factor = std::atof(e);  // tuning knob (tools/, profiles/README.md)
    int target = (int)std::lround(factor * std::cbrt((double)std::max<size_t>(leafNodes.size(), 1)));
    target = std::min(128, std::max(1, target));
    for (int k = 0; k < 3; ++k) {
        g.origin[k] = root.bmin[k];
        int d = (maxExt > 0 && ext[k] > 0) ? (int)std::ceil(target * ext[k] / maxExt) : 1;
        g.dims[k] = std::min(128, std::max(1, d));
        g.cell[k] = ext[k] > 0 ? (float)(ext[k] / g.dims[k]) : 0.f;
        g.inv_cell[k] = ext[k] > 0 ? (float)(g.dims[k] / ext[k]) : 0.f;
    }
    const double eps = 1e-3;  // outward slack in cell units: covers every float/double rounding of cell indices
    auto lo_idx = [&](double x, int k) {
        if (g.inv_cell[k] == 0.f) return 0;
        double c = std::floor((x - (double)g.origin[k]) * (double)g.inv_cell[k] - eps);
        return (int)std::min<double>(g.dims[k] - 1, std::max<double>(0, c));
    };
    auto hi_idx = [&](double x, int k) {
        if (g.inv_cell[k] == 0.f) return 0;
        double c = std::floor((x - (double)g.origin[k]) * (double)g.inv_cell[k] + eps);
        return (int)std::min<double>(g.dims[k] - 1, std::max<double>(0, c));
    };
    const size_t nCells = (size_t)g.dims[0] * g.dims[1] * g.dims[2];
    g.leaves.resize(leafNodes.size());
    for (size_t li = 0; li < leafNodes.size(); ++li) {
        const rm_bvh_node& nd = nodes[(size_t)leafNodes[li]];
        int lo[3], hi[3];
        for (int k = 0; k < 3; ++k) {
            lo[k] = lo_idx(nd.bmin[k], k);
            hi[k] = hi_idx(nd.bmax[k], k);
        }
        g.leaves[li] = {leafNodes[li], (uint32_t)(lo[0] | (lo[1] << 8) | (lo[2] << 16)), (uint32_t)(hi[0] | (hi[1] << 8) | (hi[2] << 16))};
    }
    // Every pass below scatters leaf references into per-cell lists in ascending leaf order.  The grid is cut into slabs of z
    // layers, one per thread: a thread walks ALL leaves but touches only the cells of its slab, so no two threads share a cell
    // (no atomics) and the order within a cell is the serial one.
    const int nSlabs = leafNodes.size() >= 4096 ? std::max(1, std::min({(int)std::thread::hardware_concurrency(), 16, g.dims[2]})) : 1;
    auto slabs = [&](auto&& fn) {  // fn(z0, z1): layers [z0, z1)
        if (nSlabs <= 1) return fn(0, g.dims[2]);
        std::vector<std::thread> th;
        for (int t = 0; t < nSlabs; ++t) {
            const int z0 = (int)((int64_t)g.dims[2] * t / nSlabs), z1 = (int)((int64_t)g.dims[2] * (t + 1) / nSlabs);
            if (z0 >= z1) continue;
            try {
                th.emplace_back([&fn, z0, z1] { fn(z0, z1); });
            } catch (const std::system_error&) {  // no thread to be had: this slab runs here
                fn(z0, z1);
            }
        }
        for (auto& t : th) t.join();
    };
    // cells of leaf `lr` inside layers [z0, z1); k < 0: its whole range, else only the face it is entered through by a step
    // along axis k >> 1 (k & 1: in the negative direction, i.e. the range's hi face)
    auto for_cells = [&](const LeafRef& lr, int k, int z0, int z1, auto&& fn) {
        int lo[3] = {(int)(lr.lo & 255), (int)((lr.lo >> 8) & 255), (int)((lr.lo >> 16) & 255)};
        int hi[3] = {(int)(lr.hi & 255), (int)((lr.hi >> 8) & 255), (int)((lr.hi >> 16) & 255)};
        if (k >= 0) {
            const int a = k >> 1;
            if (k & 1) lo[a] = hi[a];
            else hi[a] = lo[a];
        }
        const int zb = std::max(lo[2], z0), ze = std::min(hi[2], z1 - 1);
        for (int z = zb; z <= ze; ++z)
            for (int y = lo[1]; y <= hi[1]; ++y)
                for (int x = lo[0]; x <= hi[0]; ++x) fn(((size_t)z * g.dims[1] + y) * g.dims[0] + x);
    };
    std::vector<uint32_t> count(nCells + 1, 0u);
    slabs([&](int z0, int z1) {
        for (const LeafRef& lr : g.leaves) for_cells(lr, -1, z0, z1, [&](size_t c) { count[c + 1]++; });
    });
    for (size_t c = 0; c < nCells; ++c) count[c + 1] += count[c];
    g.cell_start = count;
    g.cell_leaf.assign(g.cell_start[nCells], 0);
    std::vector<uint32_t> fill(g.cell_start.begin(), g.cell_start.end() - 1);
    slabs([&](int z0, int z1) {
        for (size_t li = 0; li < g.leaves.size(); ++li)  // ascending leaf ordinal within each cell
            for_cells(g.leaves[li], -1, z0, z1, [&](size_t c) { g.cell_leaf[fill[c]++] = (int32_t)li; });
    });
    // ---- direction lists
    static_assert(sizeof(LeafGrid::CellDir) == 16, "CellDir is one 16-byte load");
    std::vector<uint32_t> dcount(nCells * 6, 0u);
    slabs([&](int z0, int z1) {
        for (const LeafRef& lr : g.leaves)
            for (int k = 0; k < 6; ++k) for_cells(lr, k, z0, z1, [&](size_t c) { dcount[c * 6 + k]++; });
    });
    g.cell_dir.resize(nCells);
    std::vector<uint32_t> dfill(nCells * 6, 0u);
    uint64_t total = 0;
    for (size_t c = 0; c < nCells; ++c) {
        g.cell_dir[c].base = (uint32_t)total;
        for (int k = 0; k < 6; ++k) {
            if (dcount[c * 6 + k] > 65535u) g.dir_ok = false;
            g.cell_dir[c].cnt[k] = (uint16_t)std::min(dcount[c * 6 + k], 65535u);
            dfill[c * 6 + k] = (uint32_t)total;
            total += dcount[c * 6 + k];
        }
    }
    if (total > 0xFFFFFFFFull) g.dir_ok = false;
    if (g.dir_ok) {
        g.dir_node.assign((size_t)total, 0u);
        slabs([&](int z0, int z1) {
            for (const LeafRef& lr : g.leaves)  // ascending leaf ordinal within every list
                for (int k = 0; k < 6; ++k) for_cells(lr, k, z0, z1, [&](size_t c) { g.dir_node[dfill[c * 6 + k]++] = (uint32_t)lr.node; });
        });
    } else {
        g.cell_dir.clear();
    }
}

// --------------------------------------------------------------------------------------- octree
static double box_distance(const float* amin, const float* amax, const float* bmin, const float* bmax) {  // distanceToBox
    double d[3];
    for (int k = 0; k < 3; ++k) {
        d[k] = 0;
        if (amax[k] < bmin[k]) d[k] = (double)bmin[k] - (double)amax[k];
        else if (bmax[k] < amin[k]) d[k] = (double)amin[k] - (double)bmax[k];
    }
    return hypot3(d[0], d[1], d[2]);
}

// min over every primitive box of box_distance(node box, primitive box) — what computeMinDistances (octree.ts:149-191) finds by
// trying them all, once per empty node.  A minimum does not depend on the order of its candidates, so the same double comes out
// of a pruned search: the boxes are sorted along a Morton curve and covered by an implicit binary tree of ranges; a range is
// skipped when a LOWER BOUND of the distance to its covering box already exceeds the best candidate, and every candidate that is
// looked at goes through the very same box_distance (V8's compensated hypot included).  The bound is the plain Euclidean
// gap of the covering box shrunk by 1e-9: a member box is inside the covering box, so each of its axis gaps is at least the
// cover's, and hypot3 is within a few ulp of the true norm.  Boxes with a NaN bound compare false everywhere and yield distance 0
// in the reference arithmetic; a scene that has one takes the exhaustive loop.
namespace {
class NearestBox {
public:
    explicit NearestBox(const std::vector<PrimGeom>& geom) : geom_(geom) {
        const size_t n = geom.size();
        for (const PrimGeom& g : geom)
            for (int k = 0; k < 3; ++k)
                if (std::isnan(g.bmin[k]) || std::isnan(g.bmax[k])) exhaustive_ = true;
        if (n <= 2 * kLeaf) exhaustive_ = true;
        if (exhaustive_) return;
        // Morton order of the box centres (10 bits per axis over the finite centres; infinite boxes — Repetition — sort first)
        std::vector<double> ctr(3 * n);
        for (size_t i = 0; i < n; ++i)
            for (int k = 0; k < 3; ++k) ctr[3 * i + (size_t)k] = 0.5 * ((double)geom[i].bmin[k] + (double)geom[i].bmax[k]);
        double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
        for (int k = 0; k < 3; ++k) {
            bool seeded = false;
            for (size_t i = 0; i < n; ++i) {
                const double c = ctr[3 * i + (size_t)k];
                if (!std::isfinite(c)) continue;
                lo[k] = seeded ? std::min(lo[k], c) : c;
                hi[k] = seeded ? std::max(hi[k], c) : c;
                seeded = true;
            }
        }
        std::vector<std::pair<uint32_t, int32_t>> code(n);
        for (size_t i = 0; i < n; ++i) {
            uint32_t m = 0;
            for (int k = 0; k < 3; ++k) {
                const double c = ctr[3 * i + (size_t)k], ext = hi[k] - lo[k];
                uint32_t q = 0;
                if (std::isfinite(c) && ext > 0) q = (uint32_t)std::min(1023.0, std::max(0.0, (c - lo[k]) / ext * 1024.0));
                m |= spread(q) << k;
            }
            code[i] = {m, (int32_t)i};
        }
        std::sort(code.begin(), code.end());
        order_.resize(n);
        for (size_t i = 0; i < n; ++i) order_[i] = code[i].second;
        // implicit tree over [0, n): node k covers a range, children split it in half, ranges of <= kLeaf boxes are leaves
        cover_.reserve(2 * (n / kLeaf + 1));
        build(0, n);
    }
    double min_distance(const float* amin, const float* amax) const {
        double best = std::numeric_limits<double>::infinity();
        if (exhaustive_) {
            for (const PrimGeom& g : geom_) {
                const double d = box_distance(amin, amax, g.bmin, g.bmax);
                if (d < best) best = d;
            }
            return best;
        }
        struct Item {
            uint32_t node;
            double lb;
        };
        Item stack[64];
        int sp = 0;
        stack[sp++] = {0u, 0.0};
        while (sp > 0) {
            const Item it = stack[--sp];
            if (it.lb > best) continue;
            const Cover& c = cover_[it.node];
            if (c.left == 0u) {  // leaf range
                for (uint32_t i = c.lo; i < c.hi; ++i) {
                    const PrimGeom& g = geom_[(size_t)order_[i]];
                    const double d = box_distance(amin, amax, g.bmin, g.bmax);
                    if (d < best) best = d;
                }
                continue;
            }
            const double l = lower_bound(amin, amax, cover_[c.left]), r = lower_bound(amin, amax, cover_[c.right]);
            // nearer child on top of the stack
            if (l <= r) {
                if (r <= best) stack[sp++] = {c.right, r};
                if (l <= best) stack[sp++] = {c.left, l};
            } else {
                if (l <= best) stack[sp++] = {c.left, l};
                if (r <= best) stack[sp++] = {c.right, r};
            }
        }
        return best;
    }

private:
    static constexpr size_t kLeaf = 8;
    struct Cover {
        float bmin[3], bmax[3];
        uint32_t lo, hi, left, right;  // left == 0: leaf (node 0 is the root and nobody's child)
    };
    static uint32_t spread(uint32_t v) {  // 10 bits -> every third bit
        v &= 1023u;
        v = (v | (v << 16)) & 0x030000FFu;
        v = (v | (v << 8)) & 0x0300F00Fu;
        v = (v | (v << 4)) & 0x030C30C3u;
        v = (v | (v << 2)) & 0x09249249u;
        return v;
    }
    static double lower_bound(const float* amin, const float* amax, const Cover& c) {
        double s = 0;
        for (int k = 0; k < 3; ++k) {
            double d = 0;
            if (amax[k] < c.bmin[k]) d = (double)c.bmin[k] - (double)amax[k];
            else if (c.bmax[k] < amin[k]) d = (double)amin[k] - (double)c.bmax[k];
            s += d * d;
        }
        return std::sqrt(s) * (1.0 - 1e-9);
    }
    uint32_t build(size_t lo, size_t hi) {
        const uint32_t me = (uint32_t)cover_.size();
        cover_.emplace_back();
        cover_[me].lo = (uint32_t)lo;
        cover_[me].hi = (uint32_t)hi;
        cover_[me].left = cover_[me].right = 0u;
        if (hi - lo <= kLeaf) {
            Cover& c = cover_[me];
            for (int k = 0; k < 3; ++k) {
                c.bmin[k] = std::numeric_limits<float>::infinity();
                c.bmax[k] = -std::numeric_limits<float>::infinity();
            }
            for (size_t i = lo; i < hi; ++i) {
                const PrimGeom& g = geom_[(size_t)order_[i]];
                for (int k = 0; k < 3; ++k) {
                    c.bmin[k] = std::min(c.bmin[k], g.bmin[k]);
                    c.bmax[k] = std::max(c.bmax[k], g.bmax[k]);
                }
            }
            return me;
        }
        const size_t mid = lo + (hi - lo) / 2;
        const uint32_t l = build(lo, mid), r = build(mid, hi);
        Cover& c = cover_[me];
        c.left = l;
        c.right = r;
        for (int k = 0; k < 3; ++k) {
            c.bmin[k] = std::min(cover_[l].bmin[k], cover_[r].bmin[k]);
            c.bmax[k] = std::max(cover_[l].bmax[k], cover_[r].bmax[k]);
        }
        return me;
    }
    const std::vector<PrimGeom>& geom_;
    bool exhaustive_ = false;
    std::vector<int32_t> order_;
    std::vector<Cover> cover_;
};
}  // namespace

void build_octree(const std::vector<PrimGeom>& geom, std::vector<rm_octree_node>& nodes, std::vector<int32_t>& leafPrims) {
    const int maxDepth = 6, maxPrimsPerNode = 4;  // octree.ts:39-40
    nodes.clear();
    leafPrims.clear();
    struct Pending {
        int32_t node;
        std::vector<int32_t> prims;
    };
    std::deque<Pending> q;
    nodes.emplace_back();
    {
        rm_octree_node& root = nodes[0];
        std::memset(&root, 0, sizeof(root));
        for (int k = 0; k < 3; ++k) {  // scene.ts:81-85
            root.bmin[k] = -10.f;
            root.bmax[k] = 10.f;
        }
        root.level = 0;
        std::vector<int32_t> all(geom.size());
        for (size_t i = 0; i < geom.size(); ++i) all[i] = (int32_t)i;
        q.push_back({0, std::move(all)});
    }
    while (!q.empty()) {
        Pending cur = std::move(q.front());
        q.pop_front();
        const int32_t me = cur.node;
        const int depth = nodes[(size_t)me].level;
        nodes[(size_t)me].prim_first = (int32_t)leafPrims.size();
        nodes[(size_t)me].prim_count = 0;
        nodes[(size_t)me].first_child = -1;
        nodes[(size_t)me].is_empty = 1;
        nodes[(size_t)me].min_distance = 0;
        if (depth >= maxDepth || (int)cur.prims.size() <= maxPrimsPerNode) {
            nodes[(size_t)me].prim_count = (int32_t)cur.prims.size();
            leafPrims.insert(leafPrims.end(), cur.prims.begin(), cur.prims.end());
            continue;
        }
        float bmin[3], bmax[3], c[3];
        for (int k = 0; k < 3; ++k) {
            bmin[k] = nodes[(size_t)me].bmin[k];
            bmax[k] = nodes[(size_t)me].bmax[k];
            c[k] = f32(((double)bmin[k] + (double)bmax[k]) / 2);  // BoundingBox.center (boundingBox.ts:108-114)
        }
        const int32_t first = (int32_t)nodes.size();
        nodes[(size_t)me].first_child = first;
        nodes.resize(nodes.size() + 8);
        std::vector<int32_t> childPrims[8];
        for (int i = 0; i < 8; ++i) {
            rm_octree_node& ch = nodes[(size_t)(first + i)];
            std::memset(&ch, 0, sizeof(ch));
            const int xs = i & 1, ys = (i >> 1) & 1, zs = (i >> 2) & 1;  // z,y,x nesting (octree.ts:69-88)
            ch.bmin[0] = xs ? c[0] : bmin[0];
            ch.bmax[0] = xs ? bmax[0] : c[0];
            ch.bmin[1] = ys ? c[1] : bmin[1];
            ch.bmax[1] = ys ? bmax[1] : c[1];
            ch.bmin[2] = zs ? c[2] : bmin[2];
            ch.bmax[2] = zs ? bmax[2] : c[2];
            ch.level = (uint8_t)(depth + 1);
            ch.first_child = -1;
            ch.is_empty = 1;
        }
        // closed-interval overlap of the primitive's box with every child (octree.ts:91-103).  The eight tests share their
        // per-axis halves: the low child spans [bmin, c], the high child [c, bmax] on each axis.
        for (int32_t p : cur.prims) {
            const PrimGeom& g = geom[(size_t)p];
            bool half[3][2];
            for (int k = 0; k < 3; ++k) {
                half[k][0] = bmin[k] <= g.bmax[k] && c[k] >= g.bmin[k];
                half[k][1] = c[k] <= g.bmax[k] && bmax[k] >= g.bmin[k];
            }
            for (int i = 0; i < 8; ++i)
                if (half[0][i & 1] && half[1][(i >> 1) & 1] && half[2][(i >> 2) & 1]) childPrims[i].push_back(p);
        }
        for (int i = 0; i < 8; ++i) q.push_back({first + i, std::move(childPrims[i])});
    }
    // computeMinDistances (octree.ts:149-191): bottom-up "subtree has primitives", then the
    // conservative distance of every empty node to the nearest primitive box.
    const size_t nn = nodes.size();
    std::vector<uint8_t> hasPrims(nn, 0);
    for (size_t i = nn; i-- > 0;) {
        if (nodes[i].first_child < 0) hasPrims[i] = nodes[i].prim_count > 0;
        else {
            uint8_t any = 0;
            for (int k = 0; k < 8; ++k) any |= hasPrims[(size_t)nodes[i].first_child + k];
            hasPrims[i] = any;
        }
        nodes[i].is_empty = hasPrims[i] ? 0 : 1;
    }
    std::vector<size_t> empties;
    for (size_t i = 0; i < nn; ++i)
        if (nodes[i].is_empty) empties.push_back(i);
    NearestBox nearest(geom);
    auto work = [&](size_t lo, size_t hi) {
        for (size_t e = lo; e < hi; ++e) {
            rm_octree_node& nd = nodes[empties[e]];
            const double minD = nearest.min_distance(nd.bmin, nd.bmax);
            nd.min_distance = (minD != std::numeric_limits<double>::infinity()) ? std::max(0.0, minD) : 0.0;
        }
    };
    const unsigned nt = empties.size() >= 4096 ? std::max(1u, std::min(16u, std::thread::hardware_concurrency())) : 1u;
    if (nt <= 1) {
        work(0, empties.size());
    } else {
        std::vector<std::thread> th;
        size_t per = (empties.size() + nt - 1) / nt;
        for (unsigned t = 0; t < nt; ++t) {
            size_t lo = std::min(empties.size(), t * per), hi = std::min(empties.size(), lo + per);
            if (lo >= hi) continue;
            try {
                th.emplace_back(work, lo, hi);
            } catch (const std::system_error&) {
                work(lo, hi);
            }
        }
        for (auto& t : th) t.join();
    }
}

// ------------------------------------------------------------------------------------------ operator trees
namespace {
struct NodeGeom {
    float pos[3];   // getWorldPosition()
    double radius;  // getLocalBoundingRadius()
};
double js_max(double a, double b) { return (std::isnan(a) || std::isnan(b)) ? std::numeric_limits<double>::quiet_NaN() : std::max(a, b); }

void leaf_geom(const rm_scene& s, int32_t i, bool lengthSqrt, NodeGeom& g) {
    const float* m = s.world_to_local + 16 * (size_t)i;
    float inv[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    mat4_invert(m, inv);
    g.pos[0] = inv[12];
    g.pos[1] = inv[13];
    g.pos[2] = inv[14];
    const double* q = s.params + 4 * (size_t)i;
    if (s.type[i] == RM_PRIM_SPHERE) g.radius = q[0];
    else if (s.type[i] == RM_PRIM_BOX) {
        double hx = (double)f32(q[0]), hy = (double)f32(q[1]), hz = (double)f32(q[2]);
        g.radius = lengthSqrt ? std::sqrt(hx * hx + hy * hy + hz * hz) : hypot3(hx, hy, hz);
    } else if (s.type[i] == RM_PRIM_MANDELBULB) g.radius = 2.5;  // mandelbulb.ts:80-83
    else g.radius = q[0] + q[1];
}

void node_geom(const rm_scene& s, int32_t ni, bool lengthSqrt, NodeGeom& g) {
    const rm_op_node& nd = s.op_nodes[ni];
    switch (nd.kind) {
        case RM_NODE_PRIMITIVE: leaf_geom(s, nd.prim, lengthSqrt, g); break;
        case RM_NODE_ROUND:  // round.ts:26-34
            node_geom(s, nd.child[0], lengthSqrt, g);
            g.radius = g.radius + nd.p[0];
            break;
        case RM_NODE_TWIST:               // twist.ts:38-46
        case RM_NODE_SMOOTH_SUBTRACTION:  // smoothSubstraction.ts:36-44: prim1 only
            node_geom(s, nd.child[0], lengthSqrt, g);
            break;
        case RM_NODE_REPETITION:  // repetition.ts:31-39
            node_geom(s, nd.child[0], lengthSqrt, g);
            g.radius = std::numeric_limits<double>::infinity();
            break;
        case RM_NODE_ANIMATED_TRANSLATE:  // animatedTranslate.ts:50-57
            node_geom(s, nd.child[0], lengthSqrt, g);
            g.radius = g.radius + nd.p[0];
            break;
        default: {  // smoothUnion.ts:37-59
            NodeGeom a, b;
            node_geom(s, nd.child[0], lengthSqrt, a);
            node_geom(s, nd.child[1], lengthSqrt, b);
            double dx = (double)b.pos[0] - (double)a.pos[0], dy = (double)b.pos[1] - (double)a.pos[1], dz = (double)b.pos[2] - (double)a.pos[2];
            double centerDist = lengthSqrt ? std::sqrt(dx * dx + dy * dy + dz * dz) : hypot3(dx, dy, dz);  // vec3.distance
            g.radius = js_max(a.radius, b.radius) + centerDist * 0.5;
            for (int k = 0; k < 3; ++k) g.pos[k] = f32(((double)a.pos[k] + (double)b.pos[k]) / 2);
            break;
        }
    }
}

// BoundingBox.fromPrimitive (boundingBox.ts:133-154) from (worldPos, localRadius, Primitive.transform)
void padded_box(const NodeGeom& ng, const float* m, PrimGeom& g) {
    float inv[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    const bool ok = mat4_invert(m, inv);
    const float* sm = ok ? inv : m;
    double scaleX = hypot3(sm[0], sm[1], sm[2]), scaleY = hypot3(sm[4], sm[5], sm[6]), scaleZ = hypot3(sm[8], sm[9], sm[10]);
    double maxScale = js_max(js_max(scaleX, scaleY), scaleZ);
    double r = ng.radius * maxScale * 1.5;
    for (int k = 0; k < 3; ++k) {
        g.world[k] = ng.pos[k];
        g.bmin[k] = f32((double)ng.pos[k] - r);
        g.bmax[k] = f32((double)ng.pos[k] + r);
    }
}

int check_node(const rm_scene& s, int32_t ni, int depth, std::vector<uint8_t>& onPath, std::string& err) {
    char buf[160];
    if (ni < 0 || ni >= s.n_op_nodes) {
        snprintf(buf, sizeof(buf), "operator node index %d out of range [0,%d)", ni, s.n_op_nodes);
        err = buf;
        return RM_ERR_ARG;
    }
    if (depth > RM_MAX_TREE_DEPTH) {
        snprintf(buf, sizeof(buf), "operator tree deeper than RM_MAX_TREE_DEPTH = %d", RM_MAX_TREE_DEPTH);
        err = buf;
        return RM_ERR_ARG;
    }
    if (onPath[(size_t)ni]) {
        err = "operator nodes form a cycle";
        return RM_ERR_ARG;
    }
    const rm_op_node& nd = s.op_nodes[ni];
    if (nd.kind < RM_NODE_PRIMITIVE || nd.kind > RM_NODE_ANIMATED_TRANSLATE) {
        snprintf(buf, sizeof(buf), "operator node %d has kind %d: not one of the supported SDF operators (no CPU fallback)", ni, nd.kind);
        err = buf;
        return RM_ERR_UNSUPPORTED_PRIMITIVE;
    }
    if (nd.kind == RM_NODE_PRIMITIVE) {
        if (nd.prim < 0 || nd.prim >= s.n_prims) {
            snprintf(buf, sizeof(buf), "operator node %d references primitive %d out of range [0,%d)", ni, nd.prim, s.n_prims);
            err = buf;
            return RM_ERR_ARG;
        }
        return RM_OK;
    }
    for (int k = 0; k < 16; ++k)
        if (!std::isfinite(nd.transform[k])) {
            snprintf(buf, sizeof(buf), "operator node %d: non-finite transform", ni);
            err = buf;
            return RM_ERR_ARG;
        }
    const int arity = (nd.kind == RM_NODE_SMOOTH_UNION || nd.kind == RM_NODE_SMOOTH_SUBTRACTION) ? 2 : 1;
    onPath[(size_t)ni] = 1;
    for (int k = 0; k < arity; ++k) {
        int rc = check_node(s, nd.child[k], depth + 1, onPath, err);
        if (rc) return rc;
    }
    onPath[(size_t)ni] = 0;
    return RM_OK;
}

bool is_identity(const float* m) {
    static const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    for (int k = 0; k < 16; ++k)
        if (m[k] != I[k]) return false;
    return true;
}

struct Compiler {
    const rm_scene& s;
    bool elide;
    TreeProgram& out;
    uint32_t hist[3] = {0, 0, 0};
    uint32_t flops = 0;

    void emit(int32_t op, int32_t a = 0, double c = 0.0, const float* v = nullptr) {
        DevInstr in{};
        in.op = op;
        in.a = a;
        in.c = c;
        if (v) std::memcpy(in.v, v, 3 * sizeof(float));
        out.instrs.push_back(in);
    }
    int xform(const float* m) {  // returns the number of points pushed (0 when elided)
        if (elide && is_identity(m)) return 0;
        const int32_t mi = (int32_t)(out.mats.size() / 16);
        out.mats.insert(out.mats.end(), m, m + 16);
        emit(I_XFORM, mi);
        flops += 27;
        return 1;
    }
    // `Primitive.sdf` prologue of an operator (primitive.ts:33-39) followed by its "back to world" step
    // (round.ts:17-20 etc.): local = T p, world = inv(T) local.
    int prologue(const rm_op_node& nd, bool backToWorld) {
        int pushed = xform(nd.transform);
        if (backToWorld) {
            float inv[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};  // mat4.create(): untouched when singular
            mat4_invert(nd.transform, inv);
            pushed += xform(inv);
        }
        return pushed;
    }
    // `frozen`: below an AnimatedTranslate.  Scene.updateTime reaches a node only through the setTime overrides, and
    // AnimatedTranslate's does not forward to its child — everything underneath keeps its construction-time time = 0.
    void node(int32_t ni, bool frozen = false) {
        const rm_op_node& nd = s.op_nodes[ni];
        int pushed = 0;
        switch (nd.kind) {
            case RM_NODE_PRIMITIVE:
                emit(I_PRIM, nd.prim, frozen ? 0.0 : 1.0);  // c = time scale seen by the leaf (mandelbulb.ts:33-35)
                if (s.type[nd.prim] <= RM_PRIM_TORUS) hist[s.type[nd.prim]]++;
                return;
            case RM_NODE_ROUND:
                pushed = prologue(nd, true);
                node(nd.child[0], frozen);
                if (pushed) emit(I_POP, pushed);
                emit(I_SUBC, 0, nd.p[0]);
                flops += 1;
                return;
            case RM_NODE_TWIST:
                pushed = prologue(nd, true);
                emit(I_TWIST, 0, nd.p[0]);
                flops += 9;
                node(nd.child[0], frozen);
                emit(I_POP, pushed + 1);
                return;
            case RM_NODE_REPETITION: {
                pushed = prologue(nd, true);
                const float sp[3] = {f32(nd.p[0]), f32(nd.p[1]), f32(nd.p[2])};
                emit(I_REPEAT, 0, 0.0, sp);
                flops += 12;
                node(nd.child[0], frozen);
                emit(I_POP, pushed + 1);
                return;
            }
            case RM_NODE_ANIMATED_TRANSLATE: {
                pushed = prologue(nd, false);
                AnimSlot a{};
                std::memcpy(a.dir, nd.dir, sizeof(a.dir));
                a.amplitude = nd.p[0];
                a.speed = nd.p[1];
                a.frozen = frozen;
                out.anims.push_back(a);
                emit(I_SUBV, (int32_t)out.anims.size() - 1);
                flops += 3;
                node(nd.child[0], true);
                emit(I_POP, pushed + 1);
                return;
            }
            default:
                pushed = prologue(nd, true);
                node(nd.child[0], frozen);
                node(nd.child[1], frozen);
                if (pushed) emit(I_POP, pushed);
                emit(nd.kind == RM_NODE_SMOOTH_UNION ? I_SUNION : I_SSUB, 0, nd.p[0]);
                flops += 10;
                return;
        }
    }
};
}  // namespace

int validate_tree(const rm_scene& s, std::string& err) {
    if (s.n_objects < 0 || s.n_op_nodes < 0) {
        err = "negative operator-tree counts";
        return RM_ERR_ARG;
    }
    if (s.n_objects == 0) return RM_OK;
    if (!s.op_nodes || !s.object_root || s.n_op_nodes == 0) {
        err = "n_objects > 0 needs op_nodes and object_root";
        return RM_ERR_ARG;
    }
    std::vector<uint8_t> onPath((size_t)s.n_op_nodes, 0);
    for (int32_t i = 0; i < s.n_objects; ++i) {
        int rc = check_node(s, s.object_root[i], 1, onPath, err);
        if (rc) return rc;
    }
    return RM_OK;
}

void compute_object_geometry(const rm_scene& s, unsigned flags, std::vector<PrimGeom>& out) {
    if (s.n_objects == 0) {
        compute_prim_geometry(s.n_prims, s.type, s.world_to_local, s.params, flags, out);
        return;
    }
    const bool lengthSqrt = (flags & RM_F_LENGTH_SQRT) != 0;
    out.resize((size_t)s.n_objects);
    for (int32_t i = 0; i < s.n_objects; ++i) {
        const rm_op_node& root = s.op_nodes[s.object_root[i]];
        NodeGeom ng;
        node_geom(s, s.object_root[i], lengthSqrt, ng);
        const float* m = (root.kind == RM_NODE_PRIMITIVE) ? s.world_to_local + 16 * (size_t)root.prim : root.transform;
        padded_box(ng, m, out[(size_t)i]);
    }
}

bool compile_tree(const rm_scene& s, bool elide_identity, TreeProgram& out) {
    out = TreeProgram();
    bool ok = true;
    for (int32_t i = 0; i < s.n_objects; ++i) {
        Compiler cc{s, elide_identity, out};
        out.obj_first.push_back((int32_t)out.instrs.size());
        cc.node(s.object_root[i]);
        ok = ok && cc.hist[0] <= 255u && cc.hist[1] <= 255u && cc.hist[2] <= 255u;  // packed 8-bit leaf counts
        out.obj_hist.push_back(std::min(cc.hist[0], 255u) | (std::min(cc.hist[1], 255u) << 8) | (std::min(cc.hist[2], 255u) << 16));
        out.obj_flops.push_back(cc.flops);
    }
    out.obj_first.push_back((int32_t)out.instrs.size());
    return ok;
}

void eval_anim_offsets(const TreeProgram& prog, double time, std::vector<float>& out4) {
    out4.assign(prog.anims.size() * 4, 0.f);
    for (size_t i = 0; i < prog.anims.size(); ++i) {
        const AnimSlot& a = prog.anims[i];
        const double offset = std::sin((a.frozen ? 0.0 : time) * a.speed) * a.amplitude;  // Math.sin (animatedTranslate.ts:36)
        for (int k = 0; k < 3; ++k) out4[4 * i + k] = f32((double)a.dir[k] * offset);  // vec3.scale -> f32
    }
}

}  // namespace rm
