// rm_types.h — structures shared between the host side of librm_b200.so and its sm_100a kernels.
#pragma once
#include <stdint.h>
#include <vector_types.h>  // float4 / uint3 (CUDA toolkit header, host-safe)

#include "../../include/rm.h"

#ifndef RM_PEND_CAP
#define RM_PEND_CAP 64  // pending (sorted, not yet released) intervals of the lazy grid walk; overflow -> literal list
#endif

namespace rm {

constexpr int kTileW = 8;   // image tile handed to a warp by the atomic work queue: 8 x 4 = 32 pixels
constexpr int kTileH = 4;
constexpr int kMaxStepsSphere = 100;  // sphereTracer.ts:3, adaptiveStepV2.ts:3, adaptiveStepV3.ts:3
constexpr int kMaxStepsFixed = 200;   // fixedStep.ts:3, adaptiveStep.ts:3
constexpr int kOctreeMaxDepth = 6;    // octree.ts:39
constexpr int kBvhStack = 64;         // BVH depth <= 20 (bvh.ts:32): DFS stack needs <= 2*20+1 entries

// Primitive-set specialisation chosen at upload time.
enum PrimKind {
    PK_GENERAL = 0,  // any mix of sphere/box/torus with a general affine world->local
    PK_TSPHERE = 1   // every primitive is a sphere whose world->local is a pure translation
};

// Operator trees (src/util/primitive_operations/*.ts) are compiled at upload into one linear program per scene
// object for a small stack machine: sample points only flow down the tree (point stack), distances only flow up
// (distance stack), so `Primitive.sdf` recursion becomes a straight instruction sequence.
//   Round(c, r)       : XFORM T, XFORM inv(T), <c>, POP 2, SUBC r           (round.ts:15-24)
//   Twist(c, k)       : XFORM T, XFORM inv(T), TWIST k, <c>, POP 3          (twist.ts:14-36)
//   Repetition(c, s)  : XFORM T, XFORM inv(T), REPEAT s, <c>, POP 3         (repetition.ts:14-29)
//   AnimatedTranslate : XFORM T, SUBV offset(time), <c>, POP 2              (animatedTranslate.ts:34-48)
//   SmoothUnion(a, b) : XFORM T, XFORM inv(T), <a>, <b>, POP 2, SUNION k    (smoothUnion.ts:18-35)
//   SmoothSubtraction : ... SSUB k                                          (smoothSubstraction.ts:17-34)
//   leaf primitive    : PRIM i  (Primitive.sdf of leaf i at the top point)
enum InstrOp : int32_t { I_PRIM = 0, I_XFORM, I_POP, I_TWIST, I_REPEAT, I_SUBV, I_SUBC, I_SUNION, I_SSUB };
struct DevInstr {
    int32_t op;
    int32_t a;   // PRIM: leaf index | XFORM: matrix index | POP: count | SUBV: animation slot
    double c;    // TWIST: amount | SUBC: radius | SUNION / SSUB: smoothness
    float v[4];  // REPEAT: spacing
};
constexpr int kMaxPointStack = 3 * RM_MAX_TREE_DEPTH + 2;
constexpr int kMaxDistStack = RM_MAX_TREE_DEPTH + 2;

// Fast path, translation-only spheres behind a BVH: everything the point query needs about a leaf in one 80-byte record
// indexed by BVH node index (box, up to two spheres inline with their fp64 radii) — three dependent loads per query
// (cell -> entry -> record) instead of six (cell -> entry -> node -> leaf list -> sphere -> radius).
struct LeafRecTS {
    float bmin[3];
    int32_t count;       // 1 or 2: spheres inline; 0: not a leaf; < 0: -count primitives, take the generic path through the node
    float bmax[3];
    int32_t prim_first;
    float4 s[2];         // (tx, ty, tz, r)
    double r[2];         // radii as the reference holds them (JS numbers)
};
static_assert(sizeof(LeafRecTS) == 80, "LeafRecTS layout");

// Device-resident scene (all pointers are device pointers).
struct DevScene {
    int32_t n_prims;
    int32_t accel_kind;
    int32_t prim_kind;
    int32_t n_nodes;
    // exact inputs (validation kernels read these)
    const uint8_t* type;   // [n]
    const float* w2l;      // [16n] column-major
    const double* params;  // [4n]
    // fast-path packed records
    // PK_GENERAL: 4 x float4 per primitive (3 affine rows + params/type).
    // PK_TSPHERE: structure-of-arrays per 32-primitive chunk: float tx[32], ty[32], tz[32], r[32], tt[32] = |t|^2 (640 B per chunk,
    //             the last chunk padded with far-away dummies) so that one LDS.128 feeds two packed f32x2 lanes.
    const float4* rec;
    const float4* rec1;      // PK_TSPHERE: one float4 (tx,ty,tz,r) per primitive for single-primitive evaluations
    const float* chunk_rmax;  // PK_TSPHERE: largest radius inside each 32-primitive chunk
    int32_t n_chunks;        // PK_TSPHERE: number of 32-primitive chunks
    float r_min, r_max;      // PK_TSPHERE: radius range (screening bound of the squared-distance search)
    float tt_max;            // PK_TSPHERE: largest |translation|^2 (error bound of the screening arithmetic)
    // PK_TSPHERE cluster screen on the tensor cores (rm_api.cu / tc_pass): spheres in a balanced kd order, clusters of 128.
    // tc_tiles: per 128 clusters a 4 KB shared-memory image [2 K-chunks][16 row groups][8 rows][4 tf32] of rows
    // (-C_hi, |C|^2_hi | -C_lo, |C|^2_lo) — the K-major, no-swizzle B operand of tcgen05.mma.kind::tf32.  cl_bound: (R_j, u_j)
    // per cluster.  cl_rec: sorted chunk-SoA copy of the spheres (chunk = cluster); cl_perm: sorted -> scene index.
    const float* tc_tiles;
    const float2* cl_bound;
    const float* cl_block_rmax;  // largest R_j of each 128-cluster block
    const float4* cl_rec;
    const int32_t* cl_perm;
    int32_t n_tc_blocks, n_clusters;
    // acceleration structure
    const rm_bvh_node* bvh;
    const rm_octree_node* oct;
    const int32_t* leaf_prims;
    uint32_t type_hist[3];  // number of sphere / box / torus primitives in the scene
    double time;            // Job.time of the current render (AnimatedTranslate offsets are precomputed; the Mandelbulb reads it)
    // operator trees (n_instrs > 0): n_prims counts scene OBJECTS; type / w2l / params hold the trees' leaves
    int32_t n_instrs;
    const DevInstr* instrs;
    const int32_t* obj_first;    // [n_prims + 1] program range of each object
    const float* mats;           // [16 * n_mats] column-major mat4 operands of XFORM
    const float* anim;           // [4 * n_anim] AnimatedTranslate offset vectors of the current frame (f32)
    const uint32_t* obj_hist;    // per object: leaf evaluations per call, sphere | box << 8 | torus << 16
    const uint32_t* obj_flops;   // per object: FLOPs of its operator instructions per call
    uint32_t all_op_flops;       // sum of obj_flops (one pass over every object)
    // uniform grid over the BVH leaf boxes (fast path; see rm_host.h LeafGrid)
    int32_t lazy_cap;                 // pending-interval capacity of the lazy walk (= its buffer size; smaller only under the
                                      // RM_LAZY_CAP test knob, which forces the hand-over to the literal interval list)
    const uint32_t* grid_cell_start;  // [nx*ny*nz + 1]
    const LeafRecTS* grid_leafrec;    // PK_TSPHERE: per BVH node (see LeafRecTS); null otherwise
    const uint32_t* grid_cell_node;   // per cell entry (ascending leaf ordinal within a cell): BVH node index of the leaf
    // direction lists of the DDA walk (rm_host.h LeafGrid::CellDir): x = base, y / z / w = six 16-bit counts; null = unavailable
    const uint4* grid_cell_dir;
    const uint32_t* grid_dir_node;
    int32_t grid_dims[3];
    float grid_origin[3], grid_inv[3], grid_cell[3];
};

// Per-launch statistics accumulated by the render kernel's epilogue (one atomic set per warp).
constexpr int kMaxBands = 16;  // row bands of one rm_render whose D2H is started while the kernel still runs

struct DevStats {
    unsigned long long sum_sdf, sum_iters;            // on the u16-wrapped per-pixel values
    unsigned long long sum_sdf_full, sum_iters_full;  // un-wrapped
    unsigned long long evals_sphere, evals_box, evals_torus;
    unsigned long long n_hit;
    unsigned long long op_flops;  // operator-tree scenes: FLOPs of the operator instructions executed
    unsigned long long tc_passes, tc_requests, tc_items;  // cluster screen: cooperative passes, requests served, (query, cluster) work items
    unsigned int max_sdf, min_sdf, max_iters, min_iters;
    unsigned int queue;  // atomic tile counter of the persistent-CTA work queue
    unsigned long long t_total, t_search, t_barrier, t_stuck;  // RM_PHASE_TIMING builds: warp-cycles by phase
    unsigned long long n_pass, n_req, n_rearm, t_tc[4];        // RM_PHASE_TIMING builds: cooperative passes, requests served, buffer re-arms and
                                                               // tc_pass cycles by step as seen by thread 0
    unsigned long long t_enter_min, t_drain_min, t_exit_min, t_exit_max;  // globaltimer ns: first CTA in, first empty-queue ticket, first / last warp out
    unsigned int band_done[kMaxBands];  // rm_render with page-locked planes: pixels finalised per row band (early D2H)
    unsigned int pad_;
};

struct RenderParams {
    DevScene scene;
    int32_t width, height, y_start, y_end;
    float rot3[9];
    float origin[3];
    int32_t algorithm;
    double step_size, overshoot;
    int32_t shader, shader2;
    int32_t length_sqrt;  // validation: vec3.length = sqrt(x*x+y*y+z*z) instead of Math.hypot
    int32_t anatomy;       // RM_ANATOMY=1: record the frame-anatomy timestamps (rm_stats.drain_ms / tail_ms)
    int32_t tail_trigger;  // cooperative pass at the end of the frame: parked warps that trigger it (kWarpsPerCta = only when all are)
    int32_t fast_objects;  // exact kernels in a default context (operator trees / Mandelbulb): march-step queries in fp32
    int32_t n_tiles, tiles_x;
    // cost-ordered tile queue (both optional): tile_order[t] = tile handed out with ticket t; tile_cost[tile] = cycles it took
    const unsigned int* tile_order;
    unsigned int* tile_cost;
    int32_t vec_store;  // width % 8 == 0 and every output plane 16-byte aligned: whole tiles are written with vector stores
    int32_t stripe_rows, stripe_count, stripe_index, tiles_per_stripe;  // row-stripe interleave (multi-GPU)
    // outputs (device pointers; optional ones may be null)
    uint8_t* depth;
    uint8_t* normal;
    uint16_t* sdf;
    uint16_t* iters;
    uint8_t* rgba;
    uint8_t* rgba2;
    float* depth_f32;
    uint32_t* sdf_u32;
    double* depth_f64;
    DevStats* stats;
    // early download (rm_render only; null otherwise): the thread that finalises the last pixel of a row band raises
    // band_flags[band] in page-locked host memory, and the host starts that band's D2H while the kernel still runs
    unsigned int* band_flags;
    int32_t band_rows;
    uint32_t band_px[kMaxBands];  // pixels of this launch inside each band (with row stripes: the owned rows only)
};

struct ShadeParams {
    int32_t shader;
    int32_t n_pixels;
    const uint8_t* depth;
    const uint8_t* normal;
    const uint16_t* sdf;
    const uint16_t* iters;
    uint8_t* rgba;
};

// Launchers implemented in rm_kernels_val.cu / rm_kernels_fast.cu.  Return a cudaError_t as int.
int launch_render_val(const RenderParams& p, int n_sms, void* stream);
int launch_render_fast(const RenderParams& p, int n_sms, void* stream);
int launch_shade_val(const ShadeParams& p, void* stream);
int launch_shade_fast(const ShadeParams& p, void* stream);
int probe_fp32_peak(int n_sms, void* stream, float* scratch, double* tflops);
int launch_order_tiles(const unsigned int* cost, unsigned int* order, int n_tiles, int n_runs, void* stream);  // rm_kernels_fast.cu

}  // namespace rm
