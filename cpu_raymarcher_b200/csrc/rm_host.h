// rm_host.h — host-side helpers of librm_b200.so (builders, scene packing).
#pragma once
#include <cstdint>
#include <vector>

#include <string>

#include "../../include/rm.h"
#include "rm_types.h"

namespace rm {

// Per-primitive geometry every builder needs, computed once per scene (f32 values as the reference stores them).
struct PrimGeom {
    float world[3];         // Primitive.getWorldPosition()  (primitive.ts:20-30)
    float bmin[3], bmax[3];  // BoundingBox.fromPrimitive()   (boundingBox.ts:133-154)
};

void compute_prim_geometry(int32_t n, const uint8_t* type, const float* w2l, const double* params, unsigned flags,
                           std::vector<PrimGeom>& out);
// Scene objects of an rm_scene: the primitives themselves, or operator trees over them (ABI v2).
// Geometry follows the operator overrides of getWorldPosition / getLocalBoundingRadius and uses the root
// node's own transform for BoundingBox.fromPrimitive's scale estimate.
void compute_object_geometry(const rm_scene& s, unsigned flags, std::vector<PrimGeom>& out);
// Structural check of the node arrays (indices, kinds, arity, depth <= RM_MAX_TREE_DEPTH, no sharing cycles).
// Returns RM_OK or an rm_status with a message.
int validate_tree(const rm_scene& s, std::string& err);

// One AnimatedTranslate node: offsetVec(time) = f32(direction * (sin(time * speed) * amplitude))  (animatedTranslate.ts:35-40)
struct AnimSlot {
    float dir[3];
    double amplitude, speed;
    bool frozen;  // sits under another AnimatedTranslate, whose setTime does not forward (animatedTranslate.ts:30-32): time stays 0
};
struct TreeProgram {
    std::vector<DevInstr> instrs;
    std::vector<int32_t> obj_first;   // [n_objects + 1]
    std::vector<float> mats;          // 16 floats per XFORM operand
    std::vector<AnimSlot> anims;
    std::vector<uint32_t> obj_hist, obj_flops;
};
// elide_identity: drop XFORM instructions whose matrix is exactly the identity (the smooth operators' transform):
// f32(I * p) == p for every finite f32-valued p up to the sign of zero, which no consumer can observe.
// Returns false when one object has more than 255 leaves of one primitive type.
bool compile_tree(const rm_scene& s, bool elide_identity, TreeProgram& out);
void eval_anim_offsets(const TreeProgram& prog, double time, std::vector<float>& out4);

void build_bvh(const std::vector<PrimGeom>& geom, std::vector<rm_bvh_node>& nodes, std::vector<int32_t>& leafPrims);
void build_octree(const std::vector<PrimGeom>& geom, std::vector<rm_octree_node>& nodes, std::vector<int32_t>& leafPrims);

// Uniform grid over the BVH's leaf boxes (fast path only).  It does not change any result: it answers the
// two questions the reference asks its BVH — "which leaves contain this point" (bvh.ts:95-121) and "which
// leaf boxes does this ray cross, in order of entry" (bvh.ts:126-178) — with the same boxes, just located
// through cells instead of a 17-level descent.  Cell ranges are conservative (outward epsilon).
struct LeafRef {
    int32_t node;        // index of the leaf in the rm_bvh_node array
    uint32_t lo, hi;     // packed cell range: x | y << 8 | z << 16 (inclusive)
};
struct LeafGrid {
    int32_t dims[3] = {1, 1, 1};
    float origin[3] = {0, 0, 0};   // root bmin
    float inv_cell[3] = {0, 0, 0};  // cells per world unit (0 when the root is flat on that axis)
    float cell[3] = {0, 0, 0};      // world units per cell
    std::vector<LeafRef> leaves;        // in pre-order (left-first) leaf order
    std::vector<uint32_t> cell_start;   // [nx*ny*nz + 1]
    std::vector<int32_t> cell_leaf;     // leaf ordinals, ascending within a cell
    // Direction lists of the DDA walk: entering cell c by a step along axis a in direction s, the only leaves not seen in the
    // previous cell are those whose cell range STARTS at c along that step (lo[a] == c[a] for s > 0, hi[a] == c[a] for s < 0).
    // Per cell: base offset + six counts (k = 2 * axis + (s < 0)); the lists hold BVH node indices, ascending leaf ordinal.
    struct CellDir {
        uint32_t base;
        uint16_t cnt[6];
    };
    std::vector<CellDir> cell_dir;      // [nx*ny*nz]
    std::vector<uint32_t> dir_node;
    bool dir_ok = true;                 // false: some count does not fit 16 bits (the walk then falls back to the tree)
};
void build_leaf_grid(const std::vector<rm_bvh_node>& nodes, LeafGrid& grid);

}  // namespace rm
