/* rm.h — C ABI of the B200-native raymarch hot path (librm_b200.so).
 *
 * This is the drop-in boundary for vxlerian/cpu-raymarcher's per-pixel raymarch path.  The
 * reference has no FFI; the seam its controller (src/main.ts) programs against is the Web-Worker
 * message contract of src/workers/raymarchWorker.ts:
 *     Job    (raymarchWorker.ts:10-22)  -> rm_request  (+ host-computed camera basis, see below)
 *     Result (raymarchWorker.ts:24-31)  -> rm_result   (same four arrays, same dtypes, tile-local)
 *     `new Scene(accel); scene.loadPreset(i)` (raymarchWorker.ts:37-38) -> rm_upload_scene
 *     diagnostics loop of main.ts:527-548                                -> rm_stats
 *     ShadingModel.shade (src/util/shading_models/shadingModel.ts:8-16)  -> rm_shade / rm_request.shader
 * A Node N-API addon (addon/rm_napi.cc) or any other host binds exactly these entry points; see
 * INTEGRATION.md for the reference-side stub.
 *
 * Conventions: plain pointers and sizes only; every function returns RM_OK (0) or a negative
 * rm_status; nothing throws or aborts; there is NO CPU fallback — without a CUDA device every
 * compute entry point fails with RM_ERR_CUDA.  All host pointers are borrowed for the duration of
 * the call only (rm_upload_scene copies).
 */
#ifndef RM_H
#define RM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RM_ABI_VERSION 3 /* v2: operator trees (rm_op_node, rm_scene.n_objects...), rm_build_*_scene, rm_stats_t.operator_flops
                            v3: rm_pool_* (one process drives every GPU of the box), rm_stats_t.fp32_pipe_flops / tensor_flops / n_devices */

typedef struct rm_ctx rm_ctx; /* opaque: one CUDA device, its stream, device-resident scene + frame buffers */

typedef enum rm_status {
    RM_OK = 0,
    RM_ERR_ARG = -1,                   /* null pointer, bad size, bad enum */
    RM_ERR_UNSUPPORTED_PRIMITIVE = -2, /* unknown primitive / operator kinds (no CPU fallback to route them to) */
    RM_ERR_CUDA = -3,                  /* no device, launch or copy failure (message in rm_last_error) */
    RM_ERR_STATE = -4,                 /* e.g. rm_render before rm_upload_scene */
    RM_ERR_NOMEM = -5
} rm_status;

/* rm_create flags */
#define RM_F_VALIDATE_FP64 1u /* fp64, non-fused, JS-number-exact validation kernels (bit-exact counters/hit mask) */
#define RM_F_LENGTH_SQRT 2u   /* validation only: vec3.length = sqrt(x*x+y*y+z*z) instead of Math.hypot (gl-matrix caveat) */

/* src/util/primitives/{sphere,box,torus,mandelbulb}.ts */
typedef enum rm_prim_type { RM_PRIM_SPHERE = 0, RM_PRIM_BOX = 1, RM_PRIM_TORUS = 2, RM_PRIM_MANDELBULB = 3 } rm_prim_type;
/* SDF operators (the classes under src/util/primitive_operations/).  A scene object is either a primitive or a tree of these
 * over primitives; trees are passed as a flat node array (any order, children referenced by index). */
typedef enum rm_node_kind {
    RM_NODE_PRIMITIVE = 0,          /* leaf: rm_op_node.prim indexes the scene's primitive arrays           */
    RM_NODE_ROUND = 1,              /* round.ts:15-24              p[0] = radius                            */
    RM_NODE_TWIST = 2,              /* twist.ts:14-36              p[0] = twistAmount                       */
    RM_NODE_SMOOTH_UNION = 3,       /* smoothUnion.ts:18-35        p[0] = smoothness, children prim1, prim2 */
    RM_NODE_SMOOTH_SUBTRACTION = 4, /* smoothSubstraction.ts:17-34 p[0] = smoothness, children prim1, prim2 */
    RM_NODE_REPETITION = 5,         /* repetition.ts:14-29         p[0..2] = spacing (the vec3's f32 values) */
    RM_NODE_ANIMATED_TRANSLATE = 6  /* animatedTranslate.ts:34-48  p[0] = amplitude, p[1] = speed, dir      */
} rm_node_kind;
#define RM_MAX_TREE_DEPTH 16 /* deeper operator nesting is rejected with RM_ERR_ARG */

typedef struct rm_op_node {
    int32_t kind;        /* rm_node_kind */
    int32_t child[2];    /* node indices of primitive / prim1 and prim2; -1 = none */
    int32_t prim;        /* RM_NODE_PRIMITIVE only, else -1 */
    double p[4];         /* see rm_node_kind */
    float dir[4];        /* ANIMATED_TRANSLATE: this.direction as stored (already normalised, f32); [3] unused */
    float transform[16]; /* the node's own Primitive.transform (primitive.ts:10): the wrapped primitive's for
                            Round/Twist/Repetition/AnimatedTranslate, identity for the smooth operators; ignored for leaves */
} rm_op_node; /* 128 bytes */

/* Scene.accelerationStructure: "None" | "Octree" | "BVH"  (src/util/scene.ts:20,32-36) */
typedef enum rm_accel_kind { RM_ACCEL_NONE = 0, RM_ACCEL_OCTREE = 1, RM_ACCEL_BVH = 2 } rm_accel_kind;
/* Job.algorithm (raymarchWorker.ts:49-68); unknown strings map to the sphere tracer on the host side */
typedef enum rm_algorithm {
    RM_ALG_SPHERE_TRACER = 0,    /* src/cpu_algorithms/sphereTracer.ts   */
    RM_ALG_FIXED_STEP = 1,       /* src/cpu_algorithms/fixedStep.ts      */
    RM_ALG_ADAPTIVE_STEP = 2,    /* src/cpu_algorithms/adaptiveStep.ts   */
    RM_ALG_ADAPTIVE_STEP_V2 = 3, /* src/cpu_algorithms/adaptiveStepV2.ts */
    RM_ALG_ADAPTIVE_STEP_V3 = 4  /* src/cpu_algorithms/adaptiveStepV3.ts */
} rm_algorithm;
/* createShadingModelFromValue (src/main.ts:33-45) */
typedef enum rm_shader {
    RM_SHADER_NONE = -1,
    RM_SHADER_NORMAL = 0,           /* normalModel.ts      */
    RM_SHADER_PHONG = 1,            /* phongModel.ts       */
    RM_SHADER_SDF_HEATMAP = 2,      /* SDFHeatmap.ts       */
    RM_SHADER_ITERATION_HEATMAP = 3 /* IterationHeatmap.ts */
} rm_shader;

/* Flattened BVH node (src/acceleration_structures/bvh.ts:6-22).  Pre-order, root = 0, left subtree
 * directly after its parent.  Internal nodes: prim_count = 0.  Leaves: left = right = -1 and
 * leaf_prim_index[prim_first .. prim_first+prim_count) lists scene primitive indices in the
 * reference's node.primitives order. */
typedef struct rm_bvh_node {
    float bmin[3], bmax[3];
    int32_t left, right;
    int32_t prim_first, prim_count;
} rm_bvh_node;

/* Flattened octree node (src/acceleration_structures/octree.ts:6-26).  Root = 0; the 8 children of
 * a node are consecutive at first_child .. first_child+7 in the reference's z,y,x order
 * (octree.ts:69-88); first_child = -1 for leaves. */
typedef struct rm_octree_node {
    float bmin[3], bmax[3];
    int32_t first_child;
    int32_t prim_first, prim_count;
    uint8_t level, is_empty, pad_[2];
    double min_distance;
} rm_octree_node;

/* The scene as Scene.objectSDFs + optional acceleration structure (src/util/scene.ts:12-71). */
typedef struct rm_scene {
    int32_t n_prims;
    const uint8_t* type;         /* [n_prims] rm_prim_type */
    const float* world_to_local; /* [16*n_prims] column-major mat4, exactly Primitive.transform (primitive.ts:4,10) */
    const double* params;        /* [4*n_prims] sphere: r | box: hx,hy,hz (f32-valued) | torus: R, r |
                                    mandelbulb: power, iterations, enableAnimation (0/1), animationSpeed (mandelbulb.ts:11-15) */
    int32_t accel_kind;          /* rm_accel_kind */
    /* Optional host-built structure (e.g. walked out of the reference's own BVH/Octree objects).
     * If accel_kind != NONE and nodes == NULL the library builds it natively with a builder that is
     * result-identical to bvh.ts:29-92 / octree.ts:36-118,149-191. */
    int32_t n_nodes;
    const void* nodes; /* rm_bvh_node[n_nodes] or rm_octree_node[n_nodes] */
    int32_t n_leaf_prims;
    const int32_t* leaf_prim_index;
    /* Operator trees (ABI v2).  n_objects == 0: Scene.objectSDFs = the n_prims primitives, in order.
     * n_objects > 0: Scene.objectSDFs[i] = the tree rooted at op_nodes[object_root[i]]; the primitive arrays above
     * then hold the trees' leaves, and acceleration-structure leaf lists index OBJECTS. */
    int32_t n_op_nodes;
    const rm_op_node* op_nodes;
    int32_t n_objects;
    const int32_t* object_root;
} rm_scene;

/* One frame-band request = one worker Job (raymarchWorker.ts:10-22).  The host evaluates the camera
 * (camera.ts:81-88) and passes rot3 = mat3.fromMat4(camera.getRotationMatrix()) and
 * origin = camera.getPosition() exactly as raymarcher.ts:62-67 obtains them. */
typedef struct rm_request {
    int32_t width, height;  /* full frame size */
    int32_t y_start, y_end; /* rows of this band; outputs are tile-local ((y-y_start)*width + x) */
    double time;            /* Job.time -> Scene.updateTime (raymarcher.ts:59): drives AnimatedTranslate and the Mandelbulb */
    float rot3[9];          /* column-major mat3 */
    float origin[3];
    int32_t algorithm;       /* rm_algorithm */
    double step_size;        /* FixedStep ctor (fixedStep.ts:13-16), default 0.1 */
    double overshoot_factor; /* AdaptiveStepV2/V3 ctor, default 1.2 */
    int32_t shader;           /* rm_shader for result.rgba           (RM_SHADER_NONE: skip) */
    int32_t shader_analytics; /* rm_shader for result.rgba_analytics (main.ts:506-518)      */
    /* Multi-GPU extension (0,0,0 = the plain contiguous band of the reference's partition rule,
     * main.ts:444-449).  With stripe_count > 1 the call renders only the rows y of [y_start, y_end)
     * with ((y - y_start) / stripe_rows) % stripe_count == stripe_index; outputs keep the band layout
     * ((y - y_start) * width + x), untouched rows are left alone.  stripe_rows must be a multiple of 4. */
    int32_t stripe_rows, stripe_count, stripe_index;
} rm_request;

/* Result (raymarchWorker.ts:24-31).  Buffers are CALLER-allocated.  For rm_render they are host
 * pointers; for rm_render_device they are device pointers on the context's device. */
typedef struct rm_result {
    uint8_t* depth;          /* [W*th]   Uint8ClampedArray, world units (raymarcher.ts:106)      */
    uint8_t* normal;         /* [3*W*th] Uint8ClampedArray (raymarcher.ts:103-105)               */
    uint16_t* sdf_eval;      /* [W*th]   Uint16Array, wraps mod 65536 (raymarcher.ts:119)        */
    uint16_t* iters;         /* [W*th]   Uint16Array                                             */
    uint8_t* rgba;           /* [4*W*th] optional (NULL to skip): request.shader output          */
    uint8_t* rgba_analytics; /* [4*W*th] optional: request.shader_analytics output               */
    float* depth_f32;        /* [W*th]   optional extension: unquantised rayMarch return         */
    uint32_t* sdf_eval_u32;  /* [W*th]   optional extension: un-wrapped SDF-call count           */
    double* depth_f64;       /* [W*th]   optional extension: rayMarch return as a double (exact in the validation build) */
} rm_result;

/* Diagnostics of the last completed rm_render / rm_render_device on this context. */
typedef struct rm_stats_t {
    uint64_t n_pixels;
    /* main.ts:527-548, on the u16-WRAPPED buffers (what the reference's panel shows) */
    uint64_t sum_sdf, sum_iters;
    uint32_t max_sdf, min_sdf, max_iters, min_iters;
    /* un-wrapped totals (throughput accounting: SDF evals/s) */
    uint64_t sum_sdf_full, sum_iters_full;
    uint64_t evals_by_type[3]; /* un-wrapped primitive evaluations split sphere/box/torus (mandelbulb evaluations: the rest) */
    uint64_t n_hit;            /* pixels with depth < MAX_DIST */
    double operator_flops;     /* operator-tree scenes: FLOPs of the operator nodes executed (transforms, twist, smooth min...) */
    double algorithmic_flops;  /* reference-equivalent work: every SDF call the reference counts x its FLOP count — general affine
                                  sphere 26 / box 38 / torus 29, translation-only sphere 11, 7 in the screened
                                  search of large scenes.  Equals the executed FLOPs except where executed_flops says otherwise. */
    double kernel_ms;          /* CUDA-event time of the render kernel(s) */
    double wall_ms;            /* host wall time of the whole call */
    int32_t n_launches;        /* kernels launched by the call */
    int32_t device;
    /* Translation-only spheres behind a BVH (config 4): the all-primitives fallback (scene.ts:173) is answered exactly by a
     * cluster screen on the tensor cores instead of N evaluations per request. */
    uint64_t tc_passes;        /* cooperative passes (<= 128 requests each)                                     */
    uint64_t tc_requests;      /* fallback requests served                                                      */
    uint64_t tc_items;         /* (request, 128-sphere cluster) pairs evaluated sphere by sphere                 */
    double executed_flops;     /* FLOPs actually executed for SDF work: leaf evaluations + tf32 MMAs + work items */
    /* executed_flops split by the pipe that ran them (ABI v3): the roofline numerator is fp32_pipe_flops */
    double fp32_pipe_flops;    /* FP32 (non-tensor) pipe: primitive evaluations actually performed (leaf candidates, streamed
                                  search, cluster-screen work items) x their executed FLOP count                        */
    double tensor_flops;       /* tcgen05 tf32 MMAs of the cluster screen: 2 sweeps x 128 x 128 x 16 MACs per cluster block  */
    int32_t n_devices;         /* 1, or the number of GPUs a pool call used (rm_pool_stats)                               */
    int32_t pad_;
    /* frame anatomy of the render kernel, from %globaltimer: how long the tile queue lasted, and the end-of-frame tail.
     * Only recorded when the environment has RM_ANATOMY=1 (three contended atomics per warp); 0 otherwise. */
    double drain_ms;           /* first CTA in -> the first warp finds the tile queue empty                               */
    double tail_ms;            /* ... -> the last warp leaves (pool: the slowest device's)                                */
} rm_stats_t;

/* ---- lifecycle ---------------------------------------------------------------------------- */
int rm_abi_version(void);
int rm_device_count(void); /* >= 0, or RM_ERR_CUDA */
int rm_create(rm_ctx** out, int device, unsigned flags);
void rm_destroy(rm_ctx* ctx);
const char* rm_last_error(rm_ctx* ctx); /* ctx may be NULL: error of the last failed rm_create on this thread */

/* ---- scene -------------------------------------------------------------------------------- */
int rm_upload_scene(rm_ctx* ctx, const rm_scene* scene);

/* Native builders (host side, no GPU needed).  Two-call pattern: pass nodes = NULL to get the
 * counts, then call again with buffers of that size.  Results are identical (bounds, topology,
 * leaf order) to the reference's bvh.ts / octree.ts builders on the same primitives. */
int rm_build_bvh(int32_t n_prims, const uint8_t* type, const float* world_to_local, const double* params,
                 unsigned flags /* RM_F_LENGTH_SQRT */, rm_bvh_node* nodes, int32_t* n_nodes,
                 int32_t* leaf_prim_index, int32_t* n_leaf_prims);
int rm_build_octree(int32_t n_prims, const uint8_t* type, const float* world_to_local, const double* params,
                    unsigned flags /* RM_F_LENGTH_SQRT */, rm_octree_node* nodes, int32_t* n_nodes,
                    int32_t* leaf_prim_index, int32_t* n_leaf_prims);

/* The same builders over a whole rm_scene (ABI v2): works for plain primitive lists and for operator trees, where
 * every object's getWorldPosition / getLocalBoundingRadius / transform follow the operator overrides
 * (round.ts:26-34, smoothUnion.ts:37-59, repetition.ts:31-34 ...).  scene->accel_kind / nodes are ignored. */
int rm_build_bvh_scene(const rm_scene* scene, unsigned flags, rm_bvh_node* nodes, int32_t* n_nodes, int32_t* leaf_prim_index,
                       int32_t* n_leaf_prims);
int rm_build_octree_scene(const rm_scene* scene, unsigned flags, rm_octree_node* nodes, int32_t* n_nodes,
                          int32_t* leaf_prim_index, int32_t* n_leaf_prims);

/* ---- render ------------------------------------------------------------------------------- */
int rm_render(rm_ctx* ctx, const rm_request* rq, const rm_result* host_out);
/* Same, but outputs are device pointers and nothing is copied to the host.  cuda_stream is a
 * cudaStream_t (NULL = the context's own stream).  Returns after the kernels are enqueued AND the
 * stats have been read back (it synchronises the stream). */
int rm_render_device(rm_ctx* ctx, const rm_request* rq, const rm_result* device_out, void* cuda_stream);
int rm_stats(rm_ctx* ctx, rm_stats_t* out);

/* ShadingModel.shade on existing (host) buffers: pure per-pixel map of the four quantised planes. */
int rm_shade(rm_ctx* ctx, int32_t shader, uint8_t* rgba, const uint8_t* depth, const uint8_t* normal,
             const uint16_t* sdf_eval, const uint16_t* iters, int32_t width, int32_t height);

/* Live microbenchmark of the FP32 (non-tensor) FFMA peak of this device in TFLOP/s: the roofline
 * denominator of the raymarch path.  Takes ~50 ms. */
int rm_probe_fp32_peak(rm_ctx* ctx, double* tflops);

/* Pinned (page-locked) host memory owned by the context.  Result planes that live in it are filled by direct
 * asynchronous D2H copies (no staging copy on the host); any other host pointer still works through a staging
 * buffer.  For frames of 16 MB and more the copies of finished row bands are issued while the kernel is still
 * rendering the rest of the frame (the kernel flags each completed band), so rm_render costs about the kernel
 * time.  A Node addon would back its ArrayBuffers with this memory (napi_create_external_arraybuffer). */
int rm_host_alloc(rm_ctx* ctx, size_t bytes, void** host_ptr);
int rm_host_free(rm_ctx* ctx, void* host_ptr);

/* Page-lock memory the CALLER owns (an existing ArrayBuffer, or a shared-memory frame that several processes map) so
 * that rm_render treats planes inside it like rm_host_alloc memory.  With row stripes (rm_request.stripe_count > 1)
 * rm_render writes only the rows this request owns, so one process per GPU can each call rm_render on the same
 * shared frame: every GPU downloads its own stripes over its own PCIe link, overlapped with its render, and the
 * frame is complete in host memory when the last call returns (the main thread's frame buffers, main.ts:324-329,
 * filled by N workers, main.ts:452-490).  The memory must stay mapped until rm_host_unregister / rm_destroy. */
int rm_host_register(rm_ctx* ctx, void* host_ptr, size_t bytes);
int rm_host_unregister(rm_ctx* ctx, void* host_ptr);

/* ---- multi-GPU inside ONE process: rm_pool (ABI v3) ---------------------------------------------------------
 * The reference's worker pool lives in one process: main.ts:318-321 creates the workers, main.ts:444-490 deals one row
 * band of the frame to each and awaits them all, main.ts:527-548 reduces the diagnostics over the assembled frame.  An
 * rm_pool is that pool with GPUs for workers: one rm_ctx and one host thread per device, all owned by the library.
 *   - rm_pool_upload_scene: the scene is compiled and uploaded ONCE (device 0: operator programs, BVH/octree, leaf grid,
 *     cluster data) and replicated to the other devices with peer-to-peer copies over NVLink — "each worker rebuilds the
 *     scene" (raymarchWorker.ts:37-38) becomes one build + N-1 device-to-device copies;
 *   - rm_pool_render: the band is dealt to the devices as interleaved 8-row stripes, all devices render concurrently and
 *     every device downloads its own stripes over its own PCIe link into the caller's planes while its kernel still runs
 *     (page-lock the planes with rm_pool_host_alloc / rm_pool_host_register for that);
 *   - band requests of one frame (the <= 4 concurrent Jobs of main.ts:452-486) share ONE render: the first band request of
 *     a new (camera, algorithm, parameters, size, time) key renders the whole frame across all devices into a frame cache
 *     of the pool, the others copy their rows out — callable concurrently from several host threads (N-API async workers);
 *   - rm_pool_render_device: planes in device 0's HBM; the other devices' kernels store their pixels straight into them
 *     through peer access (the tile gather of main.ts:461-468 fused into the render kernel);
 *   - rm_pool_render_frames: different frames on different devices (the Analytics rotation sweep, main.ts:438-441);
 *   - rm_pool_stats: main.ts:527-548 over the whole frame — sums, maxima and minima combined over the devices, kernel_ms =
 *     the slowest device.  The 14 words per device are already in page-locked host memory when the kernels finish, so the
 *     reduction is a host loop: a device collective would add a launch and a synchronisation per frame for 112 bytes.
 * devices == NULL or n_devices <= 0: every visible device.  A device may be listed more than once (several contexts, each with
 * its own stripe share, on one GPU — how the single-GPU test box exercises the N-way path). */
typedef struct rm_pool rm_pool;
int rm_pool_create(rm_pool** out, const int* devices, int n_devices, unsigned flags);
void rm_pool_destroy(rm_pool* pool);
const char* rm_pool_last_error(rm_pool* pool); /* pool may be NULL: error of the last failed rm_pool_create on this thread */
int rm_pool_device_count(rm_pool* pool);
int rm_pool_upload_scene(rm_pool* pool, const rm_scene* scene);
int rm_pool_render(rm_pool* pool, const rm_request* rq, const rm_result* host_out);
int rm_pool_render_device(rm_pool* pool, const rm_request* rq, const rm_result* device0_out);
/* n requests (whole frames or bands), request i rendered by device i % n_devices; host_out may be NULL (diagnostics only,
 * planes stay in device scratch) or an array of n results; stats_out (optional) receives n per-frame diagnostics. */
int rm_pool_render_frames(rm_pool* pool, const rm_request* rqs, int32_t n, const rm_result* host_out, rm_stats_t* stats_out);
int rm_pool_stats(rm_pool* pool, rm_stats_t* out);                    /* of the last rm_pool_render / _device call, whole frame */
int rm_pool_device_stats(rm_pool* pool, int32_t i, rm_stats_t* out); /* device i's share of it (load balance, per-rank kernel ms) */
int rm_pool_host_alloc(rm_pool* pool, size_t bytes, void** host_ptr);  /* page-locked for every device of the pool */
int rm_pool_host_free(rm_pool* pool, void* host_ptr);
int rm_pool_host_register(rm_pool* pool, void* host_ptr, size_t bytes);
int rm_pool_host_unregister(rm_pool* pool, void* host_ptr);
int rm_pool_alloc(rm_pool* pool, size_t bytes, void** device0_ptr);    /* device-0 memory every device of the pool can store into */
int rm_pool_free(rm_pool* pool, void* device0_ptr);
int rm_pool_memcpy_d2h(rm_pool* pool, void* host, const void* device0_ptr, size_t bytes);
int rm_pool_probe_fp32_peak(rm_pool* pool, double* tflops);           /* device 0 */

/* ---- multi-GPU plumbing (one process per GPU; see DESIGN.md "multi-GPU") -------------------- */
/* Device allocation owned by the context (freed by rm_free / rm_destroy). */
int rm_alloc(rm_ctx* ctx, size_t bytes, void** dev_ptr);
int rm_free(rm_ctx* ctx, void* dev_ptr);
/* CUDA-IPC export/import so a band rendered on rank r is stored straight into rank 0's frame over
 * NVLink by the render kernel itself (fused gather).  handle is 64 bytes. */
int rm_ipc_export(rm_ctx* ctx, void* dev_ptr, uint8_t handle[64]);
int rm_ipc_open(rm_ctx* ctx, const uint8_t handle[64], void** dev_ptr);
int rm_ipc_close(rm_ctx* ctx, void* dev_ptr);
int rm_memcpy_d2h(rm_ctx* ctx, void* host, const void* dev, size_t bytes);
int rm_memcpy_h2d(rm_ctx* ctx, void* dev, const void* host, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* RM_H */
