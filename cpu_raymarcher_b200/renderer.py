"""Thin object wrapper over the C ABI: one Context = one rm_ctx (one B200, one stream)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import ACCELS, ALGORITHMS, SHADERS, RmError


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


@dataclass
class Frame:
    """The worker `Result` (raymarchWorker.ts:24-31) plus the optional extension planes."""
    yStart: int
    yEnd: int
    depth: np.ndarray
    normal: np.ndarray
    sdfEval: np.ndarray
    iters: np.ndarray
    rgba: np.ndarray | None = None
    rgba_analytics: np.ndarray | None = None
    depth_f32: np.ndarray | None = None
    sdf_u32: np.ndarray | None = None
    depth_f64: np.ndarray | None = None


def build_bvh(types, w2l, params, flags: int = 0):
    """rm_build_bvh: native BVH builder, result-identical to bvh.ts:29-92.  No GPU needed."""
    L = _lib.lib()
    types = np.ascontiguousarray(types, np.uint8)
    w2l = np.ascontiguousarray(w2l, np.float32)
    params = np.ascontiguousarray(params, np.float64)
    # one call: a median-split BVH whose leaves are never empty has at most 2n - 1 nodes and exactly n leaf entries
    n = len(types)
    nn, nl = C.c_int32(2 * n + 1), C.c_int32(n + 1)
    nodes = (_lib.BvhNode * nn.value)()
    leaf = np.zeros(nl.value, np.int32)
    rc = L.rm_build_bvh(n, _ptr(types), _ptr(w2l), _ptr(params), flags, nodes, C.byref(nn), _ptr(leaf), C.byref(nl))
    if rc:
        raise RmError(rc, "rm_build_bvh")
    return nodes, nn.value, leaf[: nl.value]


def build_octree(types, w2l, params, flags: int = 0):
    """rm_build_octree: native octree builder, result-identical to octree.ts:36-118,149-191."""
    L = _lib.lib()
    types = np.ascontiguousarray(types, np.uint8)
    w2l = np.ascontiguousarray(w2l, np.float32)
    params = np.ascontiguousarray(params, np.float64)
    nn, nl = C.c_int32(0), C.c_int32(0)
    rc = L.rm_build_octree(len(types), _ptr(types), _ptr(w2l), _ptr(params), flags, None, C.byref(nn), None, C.byref(nl))
    if rc:
        raise RmError(rc, "rm_build_octree")
    nodes = (_lib.OctreeNode * max(nn.value, 1))()
    leaf = np.zeros(max(nl.value, 1), np.int32)
    rc = L.rm_build_octree(len(types), _ptr(types), _ptr(w2l), _ptr(params), flags, nodes, C.byref(nn), _ptr(leaf), C.byref(nl))
    if rc:
        raise RmError(rc, "rm_build_octree")
    return nodes, nn.value, leaf[: nl.value]


# numpy view of rm_op_node (include/rm.h)
OP_NODE_DTYPE = np.dtype([("kind", np.int32), ("child", np.int32, 2), ("prim", np.int32), ("p", np.float64, 4),
                          ("dir", np.float32, 4), ("transform", np.float32, 16)])
assert OP_NODE_DTYPE.itemsize == C.sizeof(_lib.OpNode) == 128


def _fill_scene(types, w2l, params, op_nodes=None, object_root=None):
    """rm_scene over numpy arrays; returns (struct, keep-alive tuple)."""
    types = np.ascontiguousarray(types, np.uint8)
    w2l = np.ascontiguousarray(w2l, np.float32)
    params = np.ascontiguousarray(params, np.float64)
    s = _lib.Scene()
    s.n_prims = len(types)
    s.type, s.world_to_local, s.params = _ptr(types), _ptr(w2l), _ptr(params)
    keep = [types, w2l, params]
    if object_root is not None and len(object_root) > 0:
        op_nodes = np.ascontiguousarray(op_nodes, OP_NODE_DTYPE)
        object_root = np.ascontiguousarray(object_root, np.int32)
        s.n_op_nodes, s.op_nodes = len(op_nodes), _ptr(op_nodes)
        s.n_objects, s.object_root = len(object_root), _ptr(object_root)
        keep += [op_nodes, object_root]
    return s, keep


def _build_scene(fn_name, node_type, types, w2l, params, op_nodes, object_root, flags):
    L = _lib.lib()
    s, keep = _fill_scene(types, w2l, params, op_nodes, object_root)
    fn = getattr(L, fn_name)
    if node_type is _lib.BvhNode:  # sizes known up front (see build_bvh): one build instead of a sizing call plus a build
        n_obj = int(s.n_objects) if s.n_objects > 0 else int(s.n_prims)
        nn, nl = C.c_int32(2 * n_obj + 1), C.c_int32(n_obj + 1)
    else:
        nn, nl = C.c_int32(0), C.c_int32(0)
        rc = fn(C.byref(s), flags, None, C.byref(nn), None, C.byref(nl))
        if rc:
            raise RmError(rc, fn_name)
    nodes = (node_type * max(nn.value, 1))()
    leaf = np.zeros(max(nl.value, 1), np.int32)
    rc = fn(C.byref(s), flags, nodes, C.byref(nn), _ptr(leaf), C.byref(nl))
    if rc:
        raise RmError(rc, fn_name)
    del keep
    return nodes, nn.value, leaf[: nl.value]


def build_bvh_scene(types, w2l, params, op_nodes=None, object_root=None, flags: int = 0):
    """rm_build_bvh_scene: the BVH over scene OBJECTS (primitives or operator trees).  No GPU needed."""
    return _build_scene("rm_build_bvh_scene", _lib.BvhNode, types, w2l, params, op_nodes, object_root, flags)


def build_octree_scene(types, w2l, params, op_nodes=None, object_root=None, flags: int = 0):
    """rm_build_octree_scene: the octree over scene OBJECTS (primitives or operator trees)."""
    return _build_scene("rm_build_octree_scene", _lib.OctreeNode, types, w2l, params, op_nodes, object_root, flags)


class Context:
    def __init__(self, device: int = 0, validate_fp64: bool = False, length_sqrt: bool = False):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        self.flags = (_lib.RM_F_VALIDATE_FP64 if validate_fp64 else 0) | (_lib.RM_F_LENGTH_SQRT if length_sqrt else 0)
        rc = self._L.rm_create(C.byref(self._h), device, self.flags)
        if rc:
            raise RmError(rc, (self._L.rm_last_error(None) or b"").decode())
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._L.rm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc:
            raise RmError(rc, (self._L.rm_last_error(self._h) or b"").decode())

    # ---- scene ----
    def upload_scene(self, types, w2l, params, accel="None", nodes=None, n_nodes=0, leaf=None, op_nodes=None,
                     object_root=None):
        """rm_upload_scene.  `op_nodes` / `object_root` (OP_NODE_DTYPE array, root indices) describe operator
        trees over the primitives; without them every primitive is a scene object."""
        s, keep = _fill_scene(types, w2l, params, op_nodes, object_root)
        types = keep[0]
        s.accel_kind = ACCELS[accel] if isinstance(accel, str) else int(accel)
        if nodes is not None:
            leaf = np.ascontiguousarray(leaf, np.int32)
            s.n_nodes = n_nodes
            s.nodes = C.cast(nodes, C.c_void_p)
            s.n_leaf_prims = len(leaf)
            s.leaf_prim_index = _ptr(leaf)
        self._check(self._L.rm_upload_scene(self._h, C.byref(s)))
        del keep
        self.n_prims = len(types)

    # ---- render ----
    @staticmethod
    def make_request(width, height, rot3, origin, algorithm="sphere-tracer", y_start=0, y_end=None, step_size=0.1,
                     overshoot=1.2, shader=None, shader_analytics=None, time=0.0, stripes=None) -> _lib.Request:
        rq = _lib.Request()
        rq.width, rq.height = width, height
        rq.y_start, rq.y_end = y_start, (height if y_end is None else y_end)
        rq.time = time
        rq.rot3 = (C.c_float * 9)(*[float(v) for v in rot3])
        rq.origin = (C.c_float * 3)(*[float(v) for v in origin])
        rq.algorithm = ALGORITHMS.get(algorithm, 0) if isinstance(algorithm, str) else int(algorithm)
        rq.step_size, rq.overshoot_factor = step_size, overshoot
        rq.shader = SHADERS[shader] if not isinstance(shader, int) else shader
        rq.shader_analytics = SHADERS[shader_analytics] if not isinstance(shader_analytics, int) else shader_analytics
        if stripes is not None:  # (rows per stripe, number of GPUs, this GPU's index)
            rq.stripe_rows, rq.stripe_count, rq.stripe_index = stripes
        return rq

    def _pinned_array(self, key: str, n: int, dtype) -> np.ndarray:
        """A numpy view over page-locked memory owned by this context (rm_host_alloc), reused per plane."""
        nbytes = max(1, n * np.dtype(dtype).itemsize)
        pool = self.__dict__.setdefault("_pinned", {})
        ent = pool.get(key)
        if ent is None or ent[1] < nbytes:
            if ent is not None:
                self._check(self._L.rm_host_free(self._h, ent[0]))
            p = C.c_void_p()
            self._check(self._L.rm_host_alloc(self._h, nbytes + nbytes // 8, C.byref(p)))
            ent = (p.value, nbytes + nbytes // 8)
            pool[key] = ent
        buf = (C.c_uint8 * nbytes).from_address(ent[0])
        return np.frombuffer(buf, dtype=dtype, count=n)

    def render(self, rq: _lib.Request, extras: bool = False, pinned: bool = False) -> Frame:
        """rm_render: host buffers out (the worker path).  pinned=True returns views over the context's page-locked
        planes (filled by direct DMA; valid until the next pinned render on this context) instead of fresh arrays."""
        th = max(0, rq.y_end - rq.y_start)
        n = th * rq.width
        new = (lambda key, cnt, dt: self._pinned_array(key, cnt, dt)) if pinned else (lambda key, cnt, dt: np.zeros(cnt, dt))
        f = Frame(rq.y_start, rq.y_end, new("depth", n, np.uint8), new("normal", 3 * n, np.uint8), new("sdf", n, np.uint16),
                  new("iters", n, np.uint16))
        res = _lib.Result()
        res.depth, res.normal, res.sdf_eval, res.iters = _ptr(f.depth), _ptr(f.normal), _ptr(f.sdfEval), _ptr(f.iters)
        if rq.shader >= 0:
            f.rgba = new("rgba", 4 * n, np.uint8)
            res.rgba = _ptr(f.rgba)
        if rq.shader_analytics >= 0:
            f.rgba_analytics = new("rgba2", 4 * n, np.uint8)
            res.rgba_analytics = _ptr(f.rgba_analytics)
        if extras:
            f.depth_f32 = new("depth_f32", n, np.float32)
            f.sdf_u32 = new("sdf_u32", n, np.uint32)
            f.depth_f64 = new("depth_f64", n, np.float64)
            res.depth_f32, res.sdf_eval_u32, res.depth_f64 = _ptr(f.depth_f32), _ptr(f.sdf_u32), _ptr(f.depth_f64)
        self._check(self._L.rm_render(self._h, C.byref(rq), C.byref(res)))
        return f

    def render_into(self, rq: _lib.Request, planes: dict) -> None:
        """rm_render into caller-owned host arrays (keys: depth, normal, sdfEval, iters and, with a shader, rgba).
        With row stripes only the rows this request owns are written, so several contexts can fill one frame."""
        res = _lib.Result()
        res.depth, res.normal = _ptr(planes["depth"]), _ptr(planes["normal"])
        res.sdf_eval, res.iters = _ptr(planes["sdfEval"]), _ptr(planes["iters"])
        if rq.shader >= 0:
            res.rgba = _ptr(planes["rgba"])
        self._check(self._L.rm_render(self._h, C.byref(rq), C.byref(res)))

    def host_register(self, arr: np.ndarray) -> None:
        """rm_host_register: page-lock caller-owned memory so rm_render DMAs straight into planes inside it."""
        self._check(self._L.rm_host_register(self._h, _ptr(arr), arr.nbytes))

    def host_unregister(self, arr: np.ndarray) -> None:
        self._check(self._L.rm_host_unregister(self._h, _ptr(arr)))

    def render_device(self, rq: _lib.Request, res: _lib.Result, stream: int | None = None):
        """rm_render_device: planes are device pointers (e.g. torch tensors' data_ptr())."""
        self._check(self._L.rm_render_device(self._h, C.byref(rq), C.byref(res), stream))

    def stats(self) -> dict:
        st = _lib.Stats()
        self._check(self._L.rm_stats(self._h, C.byref(st)))
        d = {k: getattr(st, k) for k, _ in _lib.Stats._fields_ if k != "evals_by_type"}
        d["evals_by_type"] = list(st.evals_by_type)
        return d

    def shade(self, shader, depth, normal, sdf, iters, width, height) -> np.ndarray:
        out = np.zeros(4 * width * height, np.uint8)
        sid = SHADERS[shader] if isinstance(shader, str) else int(shader)
        self._check(self._L.rm_shade(self._h, sid, _ptr(out), _ptr(np.ascontiguousarray(depth, np.uint8)),
                                     _ptr(np.ascontiguousarray(normal, np.uint8)),
                                     _ptr(np.ascontiguousarray(sdf, np.uint16)),
                                     _ptr(np.ascontiguousarray(iters, np.uint16)), width, height))
        return out

    def probe_fp32_peak(self) -> float:
        """Measured FFMA peak of this device in TFLOP/s (roofline denominator)."""
        v = C.c_double(0.0)
        self._check(self._L.rm_probe_fp32_peak(self._h, C.byref(v)))
        return v.value

    # ---- multi-GPU plumbing ----
    def alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self._L.rm_alloc(self._h, nbytes, C.byref(p)))
        return p.value

    def free(self, ptr: int):
        self._check(self._L.rm_free(self._h, ptr))

    def ipc_export(self, ptr: int) -> bytes:
        buf = (C.c_uint8 * 64)()
        self._check(self._L.rm_ipc_export(self._h, ptr, buf))
        return bytes(buf)

    def ipc_open(self, handle: bytes) -> int:
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        self._check(self._L.rm_ipc_open(self._h, buf, C.byref(p)))
        return p.value

    def ipc_close(self, ptr: int):
        self._check(self._L.rm_ipc_close(self._h, ptr))

    def memcpy_d2h(self, host: np.ndarray, dev_ptr: int):
        self._check(self._L.rm_memcpy_d2h(self._h, _ptr(host), dev_ptr, host.nbytes))

    def memcpy_h2d(self, dev_ptr: int, host: np.ndarray):
        self._check(self._L.rm_memcpy_h2d(self._h, dev_ptr, _ptr(host), host.nbytes))
