"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding over oracle/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (cpu_raymarcher_b200) must never import it.  PARITY UNPINNED: the
reference ships no tests or golden vectors; this binds a C++ restatement of its TypeScript.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ALGOS = {"sphere-tracer": 0, "fixed-step": 1, "adaptive-step": 2, "adaptive-step-v2": 3, "adaptive-step-v3": 4}
ACCELS = {"None": 0, "Octree": 1, "BVH": 2}
SHADERS = {"normal": 0, "phong": 1, "sdf-heatmap": 2, "iteration-heatmap": 3}


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle.cpp", "jsnum.hpp", "glm.hpp", "scene.hpp", "march.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, i32, u32, f64 = C.c_void_p, C.c_int, C.c_uint32, C.c_double
        L.orc_scene_new.restype = vp
        L.orc_scene_free.argtypes = [vp]
        L.orc_scene_load_preset.argtypes = [vp, i32]
        L.orc_scene_load_synthetic.argtypes = [vp, i32, u32]
        L.orc_scene_set_prims.argtypes = [vp, i32, vp, vp, vp]
        L.orc_scene_n_prims.argtypes = [vp]
        L.orc_scene_get_prims.argtypes = [vp, vp, vp, vp]
        L.orc_scene_build_accel.argtypes = [vp, i32]
        L.orc_scene_set_camera.argtypes = [vp, f64, f64]
        L.orc_scene_get_camera.argtypes = [vp, vp, vp]
        L.orc_scene_rotate_camera.argtypes = [vp, f64, f64]
        L.orc_scene_get_angles.argtypes = [vp, vp]
        L.orc_bvh_counts.argtypes = [vp, vp, vp]
        L.orc_bvh_flatten.argtypes = [vp, vp, vp, vp]
        L.orc_octree_counts.argtypes = [vp, vp, vp]
        L.orc_octree_flatten.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.orc_render.argtypes = [vp, i32, f64, f64, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, i32]
        L.orc_render_rows.argtypes = [vp, i32, f64, f64, i32, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, i32]
        L.orc_shade.argtypes = [i32, vp, vp, vp, vp, vp, i32, i32]
        L.orc_stats.argtypes = [vp, vp, C.c_size_t, vp]
        L.orc_hypot3.restype = f64
        L.orc_hypot3.argtypes = [f64, f64, f64]
        L.orc_to_u8.argtypes = [f64]
        L.orc_min2.restype = f64
        L.orc_min2.argtypes = [f64, f64]
        L.orc_max2.restype = f64
        L.orc_max2.argtypes = [f64, f64]
        L.orc_mat4_invert.argtypes = [vp, vp]
        L.orc_get_transform.argtypes = [f64, f64, f64, vp, vp]
        L.orc_prim_sdf.restype = f64
        L.orc_prim_sdf.argtypes = [vp, i32, vp]
        L.orc_scene_distance.restype = f64
        L.orc_scene_distance.argtypes = [vp, vp, vp]
        L.orc_mulberry32.restype = f64
        L.orc_mulberry32.argtypes = [u32, i32]
        L.orc_set_length_mode.argtypes = [i32]
        L.orc_scene_set_tree.argtypes = [vp, vp, vp, vp, vp, i32, vp]
        L.orc_scene_tree_counts.argtypes = [vp, vp, vp]
        L.orc_scene_get_tree.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orc_scene_set_time.argtypes = [vp, f64]
        L.orc_object_geometry.argtypes = [vp, i32, vp, vp, vp, vp]
        L.orc_js_round.restype = f64
        L.orc_js_round.argtypes = [f64]
        _LIB = L
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


@dataclass
class OracleFrame:
    depth: np.ndarray      # u8  [th*W]
    normal: np.ndarray     # u8  [th*W*3]
    sdfEval: np.ndarray    # u16 [th*W]
    iters: np.ndarray      # u16 [th*W]
    depth_f64: np.ndarray  # f64 [th*W]   unquantised rayMarch return
    sdf_full: np.ndarray   # u32 [th*W]   unwrapped
    iters_full: np.ndarray  # u32 [th*W]


class OracleScene:
    """Mirror of the reference's `Scene` as the worker uses it (raymarchWorker.ts:37-39)."""

    def __init__(self):
        self._L = lib()
        self._h = self._L.orc_scene_new()

    def __del__(self):
        try:
            self._L.orc_scene_free(self._h)
        except Exception:
            pass

    def load_preset(self, idx: int):
        if self._L.orc_scene_load_preset(self._h, idx) != 0:
            raise ValueError(f"preset {idx} is outside the hot-path scope (Mandelbulb)")
        return self

    def load_synthetic(self, n: int, seed: int = 0x5EED0001):
        self._L.orc_scene_load_synthetic(self._h, n, seed)
        return self

    def set_prims(self, types, w2l, params):
        types = np.ascontiguousarray(types, np.uint8)
        w2l = np.ascontiguousarray(w2l, np.float32).reshape(-1)
        params = np.ascontiguousarray(params, np.float64).reshape(-1)
        self._L.orc_scene_set_prims(self._h, len(types), _p(types), _p(w2l), _p(params))
        return self

    @property
    def n_prims(self):
        return self._L.orc_scene_n_prims(self._h)

    def get_prims(self):
        n = self.n_prims
        t = np.zeros(n, np.uint8)
        m = np.zeros((n, 16), np.float32)
        q = np.zeros((n, 4), np.float64)
        self._L.orc_scene_get_prims(self._h, _p(t), _p(m), _p(q))
        return t, m, q

    # ---- operator trees: flat node arrays in the layout of rm_op_node (include/rm.h)
    NODE_DTYPE = np.dtype([("kind", np.int32), ("child", np.int32, 2), ("prim", np.int32), ("p", np.float64, 4),
                           ("dir", np.float32, 4), ("transform", np.float32, 16)])

    def set_tree(self, types, w2l, params, nodes, roots):
        types = np.ascontiguousarray(types, np.uint8)
        w2l = np.ascontiguousarray(w2l, np.float32).reshape(-1)
        params = np.ascontiguousarray(params, np.float64).reshape(-1)
        nodes = np.ascontiguousarray(nodes, self.NODE_DTYPE)
        roots = np.ascontiguousarray(roots, np.int32)
        self._L.orc_scene_set_tree(self._h, _p(types), _p(w2l), _p(params), _p(nodes), len(roots), _p(roots))
        return self

    def get_tree(self):
        nn, npr = C.c_int(0), C.c_int(0)
        self._L.orc_scene_tree_counts(self._h, C.byref(nn), C.byref(npr))
        t = np.zeros(npr.value, np.uint8)
        m = np.zeros((npr.value, 16), np.float32)
        q = np.zeros((npr.value, 4), np.float64)
        nodes = np.zeros(nn.value, self.NODE_DTYPE)
        roots = np.zeros(self.n_prims, np.int32)
        self._L.orc_scene_get_tree(self._h, _p(t), _p(m), _p(q), _p(nodes), _p(roots))
        return t, m, q, nodes, roots

    def set_time(self, time: float):
        """Scene.updateTime (scene.ts:135-140), which raymarcher.ts:59 calls once per job."""
        self._L.orc_scene_set_time(self._h, float(time))
        return self

    def object_geometry(self, i: int):
        w = np.zeros(3, np.float32)
        r = C.c_double(0)
        bmin = np.zeros(3, np.float32)
        bmax = np.zeros(3, np.float32)
        self._L.orc_object_geometry(self._h, i, _p(w), C.byref(r), _p(bmin), _p(bmax))
        return w, r.value, bmin, bmax

    def build_accel(self, kind):
        self._L.orc_scene_build_accel(self._h, ACCELS[kind] if isinstance(kind, str) else kind)
        return self

    def set_camera(self, pitch: float, yaw: float):
        self._L.orc_scene_set_camera(self._h, pitch, yaw)
        return self

    def rotate_camera(self, dpitch: float, dyaw: float):
        self._L.orc_scene_rotate_camera(self._h, dpitch, dyaw)
        return self

    def get_angles(self):
        a = np.zeros(2, np.float64)
        self._L.orc_scene_get_angles(self._h, _p(a))
        return float(a[0]), float(a[1])

    def get_camera(self):
        r = np.zeros(9, np.float32)
        o = np.zeros(3, np.float32)
        self._L.orc_scene_get_camera(self._h, _p(r), _p(o))
        return r, o

    def bvh_flat(self):
        nn, npr = C.c_int(0), C.c_int(0)
        if self._L.orc_bvh_counts(self._h, C.byref(nn), C.byref(npr)) != 0:
            raise RuntimeError("no BVH built")
        bounds = np.zeros((nn.value, 6), np.float32)
        links = np.zeros((nn.value, 4), np.int32)
        leaf = np.zeros(max(npr.value, 1), np.int32)
        self._L.orc_bvh_flatten(self._h, _p(bounds), _p(links), _p(leaf))
        return bounds, links, leaf[: npr.value]

    def octree_flat(self):
        nn, npr = C.c_int(0), C.c_int(0)
        if self._L.orc_octree_counts(self._h, C.byref(nn), C.byref(npr)) != 0:
            raise RuntimeError("no octree built")
        n = nn.value
        bounds = np.zeros((n, 6), np.float32)
        links = np.zeros((n, 3), np.int32)
        level = np.zeros(n, np.uint8)
        empty = np.zeros(n, np.uint8)
        mind = np.zeros(n, np.float64)
        leaf = np.zeros(max(npr.value, 1), np.int32)
        self._L.orc_octree_flatten(self._h, _p(bounds), _p(links), _p(level), _p(empty), _p(mind), _p(leaf))
        return bounds, links, level, empty, mind, leaf[: npr.value]

    def render(self, width, height, algorithm="sphere-tracer", y_start=0, y_end=None, step_size=0.1,
               overshoot=1.2, nthreads=None) -> OracleFrame:
        y_end = height if y_end is None else y_end
        th = max(0, y_end - y_start)
        n = th * width
        f = OracleFrame(np.zeros(n, np.uint8), np.zeros(n * 3, np.uint8), np.zeros(n, np.uint16),
                        np.zeros(n, np.uint16), np.zeros(n, np.float64), np.zeros(n, np.uint32),
                        np.zeros(n, np.uint32))
        if n == 0:
            return f
        nthreads = nthreads or (os.cpu_count() or 1)
        algo = ALGOS.get(algorithm, 0)  # unknown -> sphere tracer (raymarchWorker.ts:66-67)
        self._L.orc_render(self._h, algo, step_size, overshoot, width, height, y_start, y_end, _p(f.depth),
                           _p(f.normal), _p(f.sdfEval), _p(f.iters), _p(f.depth_f64), _p(f.sdf_full),
                           _p(f.iters_full), nthreads)
        return f

    def render_rows(self, width, height, rows, algorithm="sphere-tracer", step_size=0.1, overshoot=1.2,
                    nthreads=None) -> OracleFrame:
        """Bounded sample: only image rows `rows` (compact output, row k = image row rows[k])."""
        rows = np.ascontiguousarray(rows, np.int32)
        n = len(rows) * width
        f = OracleFrame(np.zeros(n, np.uint8), np.zeros(n * 3, np.uint8), np.zeros(n, np.uint16),
                        np.zeros(n, np.uint16), np.zeros(n, np.float64), np.zeros(n, np.uint32),
                        np.zeros(n, np.uint32))
        nthreads = nthreads or (os.cpu_count() or 1)
        self._L.orc_render_rows(self._h, ALGOS.get(algorithm, 0), step_size, overshoot, width, height, _p(rows),
                                len(rows), _p(f.depth), _p(f.normal), _p(f.sdfEval), _p(f.iters), _p(f.depth_f64),
                                _p(f.sdf_full), _p(f.iters_full), nthreads)
        return f

    def prim_sdf(self, prim: int, p) -> float:
        p = np.ascontiguousarray(p, np.float32)
        return self._L.orc_prim_sdf(self._h, prim, _p(p))

    def distance(self, p):
        p = np.ascontiguousarray(p, np.float32)
        c = C.c_uint32(0)
        d = self._L.orc_scene_distance(self._h, _p(p), C.byref(c))
        return d, c.value


def shade(shader, depth, normal, sdf, iters, w, h) -> np.ndarray:
    out = np.zeros(w * h * 4, np.uint8)
    sid = SHADERS[shader] if isinstance(shader, str) else shader
    lib().orc_shade(sid, _p(out), _p(depth), _p(normal), _p(sdf), _p(iters), w, h)
    return out


def stats(sdf, iters):
    out = np.zeros(4, np.float64)
    lib().orc_stats(_p(sdf), _p(iters), sdf.size, _p(out))
    return dict(total_sdf=out[0], max_sdf=out[1], min_sdf=out[2], total_iters=out[3])
