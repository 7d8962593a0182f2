// leafgrid_harness.cc — TEST INFRASTRUCTURE: unit test of rm::build_leaf_grid (csrc/rm_build.cpp), the uniform grid over the BVH's
// leaf boxes that the fast path walks instead of descending the tree (bvh.ts:95-178 asks the same two questions of the same boxes).
// The product fills the per-cell lists with z-slab threads; this harness rebuilds every list the slow, obvious way — for each cell,
// scan all leaves — and demands identical arrays, plus the containment property the kernel relies on: every leaf box lies inside
// the union of the cells it is listed in.
//
//   leafgrid_harness <n_spheres> <r_lo> <r_hi> <seed> [flat]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../cpu_raymarcher_b200/csrc/rm_host.h"

using namespace rm;

static int fail(const char* what, long a = 0, long b = 0) {
    std::fprintf(stderr, "leafgrid_harness: %s (%ld, %ld)\n", what, a, b);
    return 1;
}

int main(int argc, char** argv) {
    if (argc < 5) return fail("usage: leafgrid_harness n r_lo r_hi seed [flat]");
    const int n = std::atoi(argv[1]);
    const float rlo = (float)std::atof(argv[2]), rhi = (float)std::atof(argv[3]);
    const bool flat = argc > 5;
    std::mt19937 rng((unsigned)std::atoi(argv[4]));
    std::uniform_real_distribution<float> U(-4.f, 4.f), R(rlo, rhi);
    std::vector<uint8_t> type((size_t)n, 0);
    std::vector<float> w2l((size_t)n * 16, 0.f);
    std::vector<double> params((size_t)n * 4, 0.0);
    for (int i = 0; i < n; ++i) {
        float* m = &w2l[(size_t)i * 16];
        m[0] = m[5] = m[10] = m[15] = 1.f;
        m[12] = U(rng);
        m[13] = U(rng);
        m[14] = flat ? 0.f : U(rng);
        params[(size_t)i * 4] = R(rng);
    }
    std::vector<PrimGeom> geom;
    compute_prim_geometry(n, type.data(), w2l.data(), params.data(), 0, geom);
    std::vector<rm_bvh_node> bvh;
    std::vector<int32_t> leafPrims;
    build_bvh(geom, bvh, leafPrims);
    LeafGrid g;
    build_leaf_grid(bvh, g);

    const int nx = g.dims[0], ny = g.dims[1], nz = g.dims[2];
    const size_t nCells = (size_t)nx * ny * nz;
    if (g.cell_start.size() != nCells + 1 || g.cell_start[0] != 0 || g.cell_start[nCells] != g.cell_leaf.size()) return fail("cell_start shape");
    // leaves: exactly the BVH's non-empty leaf nodes, in node order
    size_t li = 0;
    for (size_t i = 0; i < bvh.size(); ++i)
        if (bvh[i].left < 0 && bvh[i].right < 0 && bvh[i].prim_count > 0) {
            if (li >= g.leaves.size() || g.leaves[li].node != (int32_t)i) return fail("leaf order", (long)li, (long)i);
            ++li;
        }
    if (li != g.leaves.size()) return fail("leaf count", (long)li, (long)g.leaves.size());
    auto unpack = [](uint32_t v, int* o) {
        o[0] = (int)(v & 255);
        o[1] = (int)((v >> 8) & 255);
        o[2] = (int)((v >> 16) & 255);
    };
    // containment: the cell range of a leaf covers its box (cells are [origin + i * cell, origin + (i + 1) * cell))
    for (const LeafRef& lr : g.leaves) {
        int lo[3], hi[3];
        unpack(lr.lo, lo);
        unpack(lr.hi, hi);
        const rm_bvh_node& nd = bvh[(size_t)lr.node];
        for (int k = 0; k < 3; ++k) {
            if (lo[k] > hi[k] || hi[k] >= g.dims[k]) return fail("cell range", lo[k], hi[k]);
            if (g.inv_cell[k] == 0.f) continue;
            const double cl = (double)g.origin[k] + (double)lo[k] * (double)g.cell[k], ch = (double)g.origin[k] + (double)(hi[k] + 1) * (double)g.cell[k];
            const double slack = 1e-4 * (double)g.cell[k];
            if (lo[k] > 0 && (double)nd.bmin[k] < cl - slack) return fail("box sticks out below its cells", lr.node, k);
            if (hi[k] < g.dims[k] - 1 && (double)nd.bmax[k] > ch + slack) return fail("box sticks out above its cells", lr.node, k);
        }
    }
    // per-cell lists and direction lists, the obvious way
    size_t dirTotal = 0;
    for (size_t c = 0; c < nCells; ++c) {
        const int x = (int)(c % (size_t)nx), y = (int)((c / (size_t)nx) % (size_t)ny), z = (int)(c / ((size_t)nx * ny));
        const int cc[3] = {x, y, z};
        uint32_t at = g.cell_start[c];
        uint32_t dAt[6];
        if (g.dir_ok) {
            uint32_t b = g.cell_dir[c].base;
            if (b != dirTotal) return fail("direction-list base", (long)c, (long)b);
            for (int k = 0; k < 6; ++k) {
                dAt[k] = b;
                b += g.cell_dir[c].cnt[k];
            }
            dirTotal = b;
        }
        for (size_t l = 0; l < g.leaves.size(); ++l) {
            int lo[3], hi[3];
            unpack(g.leaves[l].lo, lo);
            unpack(g.leaves[l].hi, hi);
            bool in = true;
            for (int k = 0; k < 3; ++k) in = in && lo[k] <= cc[k] && cc[k] <= hi[k];
            if (!in) continue;
            if (at >= g.cell_start[c + 1] || g.cell_leaf[at] != (int32_t)l) return fail("cell list", (long)c, (long)l);
            ++at;
            if (!g.dir_ok) continue;
            for (int k = 0; k < 6; ++k) {  // entered by a step along axis k >> 1: new exactly when the range starts at this cell on that side
                const int a = k >> 1;
                const bool starts = (k & 1) ? hi[a] == cc[a] : lo[a] == cc[a];
                if (!starts) continue;
                if (dAt[k] >= g.dir_node.size() || g.dir_node[dAt[k]] != (uint32_t)g.leaves[l].node) return fail("direction list", (long)c, k);
                ++dAt[k];
            }
        }
        if (at != g.cell_start[c + 1]) return fail("cell list too long", (long)c);
        if (g.dir_ok) {
            uint32_t b = g.cell_dir[c].base;
            for (int k = 0; k < 6; ++k) {
                b += g.cell_dir[c].cnt[k];
                if (dAt[k] != b) return fail("direction list length", (long)c, k);
            }
        }
    }
    if (g.dir_ok && dirTotal != g.dir_node.size()) return fail("direction-list total", (long)dirTotal, (long)g.dir_node.size());
    std::printf("OK leaves=%zu cells=%dx%dx%d refs=%zu dir=%zu\n", g.leaves.size(), nx, ny, nz, g.cell_leaf.size(), g.dir_node.size());
    return 0;
}
