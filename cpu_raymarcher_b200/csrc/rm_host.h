// rm_host.h — host-side helpers of librm_b200.so (builders, scene packing).
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/rm.h"

namespace rm {

// Per-primitive geometry every builder needs, computed once per scene (f32 values as the reference stores them).
struct PrimGeom {
    float world[3];         // Primitive.getWorldPosition()  (primitive.ts:20-30)
    float bmin[3], bmax[3];  // BoundingBox.fromPrimitive()   (boundingBox.ts:133-154)
};

void compute_prim_geometry(int32_t n, const uint8_t* type, const float* w2l, const double* params, unsigned flags,
                           std::vector<PrimGeom>& out);
void build_bvh(const std::vector<PrimGeom>& geom, std::vector<rm_bvh_node>& nodes, std::vector<int32_t>& leafPrims);
void build_octree(const std::vector<PrimGeom>& geom, std::vector<rm_octree_node>& nodes, std::vector<int32_t>& leafPrims);

}  // namespace rm
