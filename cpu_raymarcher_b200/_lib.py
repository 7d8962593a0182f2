"""ctypes binding of librm_b200.so (include/rm.h).  The library is the product; if it is missing this
module raises — there is no Python/CPU fallback for the raymarch path."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librm_b200.so")
if os.environ.get("RM_B200_LIB"):  # A/B experiments (tools/ab_bench.py): another build of the SAME library, e.g. build/exp/librm_b200_<variant>.so
    LIB_PATH = os.path.abspath(os.environ["RM_B200_LIB"])

RM_OK = 0
RM_ERR_ARG, RM_ERR_UNSUPPORTED_PRIMITIVE, RM_ERR_CUDA, RM_ERR_STATE, RM_ERR_NOMEM = -1, -2, -3, -4, -5
RM_F_VALIDATE_FP64, RM_F_LENGTH_SQRT = 1, 2
STATUS_NAMES = {0: "RM_OK", -1: "RM_ERR_ARG", -2: "RM_ERR_UNSUPPORTED_PRIMITIVE", -3: "RM_ERR_CUDA",
                -4: "RM_ERR_STATE", -5: "RM_ERR_NOMEM"}

ALGORITHMS = {"sphere-tracer": 0, "fixed-step": 1, "adaptive-step": 2, "adaptive-step-v2": 3, "adaptive-step-v3": 4}
ACCELS = {"None": 0, "Octree": 1, "BVH": 2}
SHADERS = {None: -1, "none": -1, "normal": 0, "phong": 1, "sdf-heatmap": 2, "iteration-heatmap": 3}


class RmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{STATUS_NAMES.get(code, code)}: {msg}")
        self.code = code


class BvhNode(C.Structure):
    _fields_ = [("bmin", C.c_float * 3), ("bmax", C.c_float * 3), ("left", C.c_int32), ("right", C.c_int32),
                ("prim_first", C.c_int32), ("prim_count", C.c_int32)]


class OctreeNode(C.Structure):
    _fields_ = [("bmin", C.c_float * 3), ("bmax", C.c_float * 3), ("first_child", C.c_int32), ("prim_first", C.c_int32),
                ("prim_count", C.c_int32), ("level", C.c_uint8), ("is_empty", C.c_uint8), ("pad_", C.c_uint8 * 2),
                ("min_distance", C.c_double)]


class OpNode(C.Structure):  # rm_op_node, 128 bytes
    _fields_ = [("kind", C.c_int32), ("child", C.c_int32 * 2), ("prim", C.c_int32), ("p", C.c_double * 4),
                ("dir", C.c_float * 4), ("transform", C.c_float * 16)]


NODE_KINDS = {"primitive": 0, "round": 1, "twist": 2, "smooth-union": 3, "smooth-subtraction": 4, "repetition": 5,
              "animated-translate": 6}
RM_MAX_TREE_DEPTH = 16


class Scene(C.Structure):
    _fields_ = [("n_prims", C.c_int32), ("type", C.c_void_p), ("world_to_local", C.c_void_p), ("params", C.c_void_p),
                ("accel_kind", C.c_int32), ("n_nodes", C.c_int32), ("nodes", C.c_void_p), ("n_leaf_prims", C.c_int32),
                ("leaf_prim_index", C.c_void_p), ("n_op_nodes", C.c_int32), ("op_nodes", C.c_void_p),
                ("n_objects", C.c_int32), ("object_root", C.c_void_p)]


class Request(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("y_start", C.c_int32), ("y_end", C.c_int32),
                ("time", C.c_double), ("rot3", C.c_float * 9), ("origin", C.c_float * 3), ("algorithm", C.c_int32),
                ("step_size", C.c_double), ("overshoot_factor", C.c_double), ("shader", C.c_int32),
                ("shader_analytics", C.c_int32), ("stripe_rows", C.c_int32), ("stripe_count", C.c_int32),
                ("stripe_index", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("depth", C.c_void_p), ("normal", C.c_void_p), ("sdf_eval", C.c_void_p), ("iters", C.c_void_p),
                ("rgba", C.c_void_p), ("rgba_analytics", C.c_void_p), ("depth_f32", C.c_void_p),
                ("sdf_eval_u32", C.c_void_p), ("depth_f64", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("n_pixels", C.c_uint64), ("sum_sdf", C.c_uint64), ("sum_iters", C.c_uint64), ("max_sdf", C.c_uint32),
                ("min_sdf", C.c_uint32), ("max_iters", C.c_uint32), ("min_iters", C.c_uint32),
                ("sum_sdf_full", C.c_uint64), ("sum_iters_full", C.c_uint64), ("evals_by_type", C.c_uint64 * 3),
                ("n_hit", C.c_uint64), ("operator_flops", C.c_double), ("algorithmic_flops", C.c_double), ("kernel_ms", C.c_double), ("wall_ms", C.c_double), ("n_launches", C.c_int32),
                ("device", C.c_int32), ("tc_passes", C.c_uint64), ("tc_requests", C.c_uint64), ("tc_items", C.c_uint64),
                ("executed_flops", C.c_double), ("fp32_pipe_flops", C.c_double), ("tensor_flops", C.c_double),
                ("n_devices", C.c_int32), ("pad_", C.c_int32), ("drain_ms", C.c_double), ("tail_ms", C.c_double)]


# every symbol include/rm.h declares
EXPORTS = ["rm_abi_version", "rm_device_count", "rm_create", "rm_destroy", "rm_last_error", "rm_upload_scene",
           "rm_build_bvh", "rm_build_octree", "rm_build_bvh_scene", "rm_build_octree_scene", "rm_render", "rm_render_device", "rm_stats", "rm_shade", "rm_alloc",
           "rm_free", "rm_host_alloc", "rm_host_free", "rm_host_register", "rm_host_unregister", "rm_probe_fp32_peak", "rm_ipc_export", "rm_ipc_open", "rm_ipc_close", "rm_memcpy_d2h", "rm_memcpy_h2d",
           "rm_pool_create", "rm_pool_destroy", "rm_pool_last_error", "rm_pool_device_count", "rm_pool_upload_scene", "rm_pool_render",
           "rm_pool_render_device", "rm_pool_render_frames", "rm_pool_stats", "rm_pool_device_stats", "rm_pool_host_alloc", "rm_pool_host_free",
           "rm_pool_host_register", "rm_pool_host_unregister", "rm_pool_alloc", "rm_pool_free", "rm_pool_memcpy_d2h", "rm_pool_probe_fp32_peak"]

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C cpu_raymarcher_b200/csrc). There is no CPU fallback for the raymarch path.")
        L = C.CDLL(LIB_PATH)
        vp, i32, u32 = C.c_void_p, C.c_int32, C.c_uint
        L.rm_abi_version.restype = C.c_int
        L.rm_device_count.restype = C.c_int
        L.rm_create.argtypes = [C.POINTER(vp), C.c_int, u32]
        L.rm_destroy.argtypes = [vp]
        L.rm_destroy.restype = None
        L.rm_last_error.argtypes = [vp]
        L.rm_last_error.restype = C.c_char_p
        L.rm_upload_scene.argtypes = [vp, C.POINTER(Scene)]
        L.rm_build_bvh.argtypes = [i32, vp, vp, vp, u32, vp, C.POINTER(i32), vp, C.POINTER(i32)]
        L.rm_build_octree.argtypes = [i32, vp, vp, vp, u32, vp, C.POINTER(i32), vp, C.POINTER(i32)]
        L.rm_build_bvh_scene.argtypes = [C.POINTER(Scene), u32, vp, C.POINTER(i32), vp, C.POINTER(i32)]
        L.rm_build_octree_scene.argtypes = [C.POINTER(Scene), u32, vp, C.POINTER(i32), vp, C.POINTER(i32)]
        L.rm_render.argtypes = [vp, C.POINTER(Request), C.POINTER(Result)]
        L.rm_render_device.argtypes = [vp, C.POINTER(Request), C.POINTER(Result), vp]
        L.rm_stats.argtypes = [vp, C.POINTER(Stats)]
        L.rm_shade.argtypes = [vp, i32, vp, vp, vp, vp, vp, i32, i32]
        L.rm_probe_fp32_peak.argtypes = [vp, C.POINTER(C.c_double)]
        L.rm_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
        L.rm_host_free.argtypes = [vp, vp]
        L.rm_host_register.argtypes = [vp, vp, C.c_size_t]
        L.rm_host_unregister.argtypes = [vp, vp]
        L.rm_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
        L.rm_free.argtypes = [vp, vp]
        L.rm_ipc_export.argtypes = [vp, vp, vp]
        L.rm_ipc_open.argtypes = [vp, vp, C.POINTER(vp)]
        L.rm_ipc_close.argtypes = [vp, vp]
        L.rm_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
        L.rm_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
        if os.environ.get("RM_B200_LIB") and not hasattr(L, "rm_pool_create"):
            _LIB = L  # an older A/B build of the library (tools/ab_bench.py) without the rm_pool entry points
            return _LIB
        L.rm_pool_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int, u32]
        L.rm_pool_destroy.argtypes = [vp]
        L.rm_pool_destroy.restype = None
        L.rm_pool_last_error.argtypes = [vp]
        L.rm_pool_last_error.restype = C.c_char_p
        L.rm_pool_device_count.argtypes = [vp]
        L.rm_pool_upload_scene.argtypes = [vp, C.POINTER(Scene)]
        L.rm_pool_render.argtypes = [vp, C.POINTER(Request), C.POINTER(Result)]
        L.rm_pool_render_device.argtypes = [vp, C.POINTER(Request), C.POINTER(Result)]
        L.rm_pool_render_frames.argtypes = [vp, C.POINTER(Request), i32, C.POINTER(Result), C.POINTER(Stats)]
        L.rm_pool_stats.argtypes = [vp, C.POINTER(Stats)]
        L.rm_pool_device_stats.argtypes = [vp, i32, C.POINTER(Stats)]
        L.rm_pool_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
        L.rm_pool_host_free.argtypes = [vp, vp]
        L.rm_pool_host_register.argtypes = [vp, vp, C.c_size_t]
        L.rm_pool_host_unregister.argtypes = [vp, vp]
        L.rm_pool_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
        L.rm_pool_free.argtypes = [vp, vp]
        L.rm_pool_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
        L.rm_pool_probe_fp32_peak.argtypes = [vp, C.POINTER(C.c_double)]
        _LIB = L
    return _LIB
