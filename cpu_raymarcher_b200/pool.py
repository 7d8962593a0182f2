"""RaymarchPool — the reference's worker POOL (src/main.ts:318-321,444-490) served by every GPU of the box from ONE process.

Thin ctypes wrapper over the rm_pool_* entry points of librm_b200.so: the library owns one context and one host thread per
device, replicates the scene device-to-device, deals interleaved row stripes to the devices, lets every device download its own
stripes into the caller's page-locked planes, and reduces the diagnostics (main.ts:527-548).  `on_message(job)` takes the same
Job the workers get (raymarchWorker.ts:10-22) — a whole frame or one of the <= 4 row bands of a frame; band jobs of one frame
share a single render (the pool's frame cache).  No torch, no torchrun, no NCCL process group on this path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import RmError
from .renderer import Context, Frame, _fill_scene, _ptr
from .scene import Scene


def _stats_dict(st: _lib.Stats) -> dict:
    d = {k: getattr(st, k) for k, _ in _lib.Stats._fields_ if k not in ("evals_by_type", "pad_")}
    d["evals_by_type"] = list(st.evals_by_type)
    return d


class RaymarchPool:
    def __init__(self, devices=None, validate_fp64: bool = False, length_sqrt: bool = False):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        self.flags = (_lib.RM_F_VALIDATE_FP64 if validate_fp64 else 0) | (_lib.RM_F_LENGTH_SQRT if length_sqrt else 0)
        if devices is None:
            rc = self._L.rm_pool_create(C.byref(self._h), None, 0, self.flags)
        else:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = self._L.rm_pool_create(C.byref(self._h), arr, len(devices), self.flags)
        if rc:
            raise RmError(rc, (self._L.rm_pool_last_error(None) or b"").decode())
        self.n_devices = self._L.rm_pool_device_count(self._h)
        self._scene_key = None
        self.scene: Scene | None = None
        self.n_prims = 0
        self._pinned = {}
        self._dev_frame = None  # (ptr, bytes) of the device-0 planes of render_device

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_h", None):
            self._L.rm_pool_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc:
            raise RmError(rc, (self._L.rm_pool_last_error(self._h) or b"").decode())

    # ------------------------------------------------------------------ scene
    def upload_scene(self, types, w2l, params, accel="None", op_nodes=None, object_root=None):
        s, keep = _fill_scene(types, w2l, params, op_nodes, object_root)
        s.accel_kind = _lib.ACCELS[accel] if isinstance(accel, str) else int(accel)
        self._check(self._L.rm_pool_upload_scene(self._h, C.byref(s)))
        self.n_prims = len(keep[0])
        del keep

    def _ensure_scene(self, preset_index: int, accel: str, synthetic=None):
        accel = accel if accel in ("Octree", "BVH") else "None"
        key = (preset_index, synthetic, accel)
        if key != self._scene_key:
            sc = Scene(accel)
            if synthetic is not None:
                sc.load_synthetic(*synthetic)
            else:
                sc.load_preset(preset_index)
            t, m, q = sc.primitives.arrays()
            self.upload_scene(t, m, q, accel, op_nodes=sc.primitives.op_nodes, object_root=sc.primitives.object_root)
            self.scene = sc
            self._scene_key = key
        return self.scene

    def _request(self, job: dict, shader, shader_analytics=None) -> _lib.Request:
        width, height = int(job["width"]), int(job["height"])
        scene = self._ensure_scene(int(job.get("scenePresetIndex", 0)), job.get("accelerationStructure", "None"), job.get("synthetic"))
        cam = job.get("camera", {})
        scene.camera.set_angles(float(cam.get("pitch", 0.0)), float(cam.get("yaw", 0.0)))
        return Context.make_request(
            width, height, scene.camera.get_rotation_matrix3(), scene.camera.get_position(), algorithm=job.get("algorithm", "sphere-tracer"),
            y_start=int(job.get("yStart", 0)), y_end=int(job.get("yEnd", height)),
            step_size=float(job["stepSize"]) if job.get("stepSize") is not None else 0.1,
            overshoot=float(job["overshootFactor"]) if job.get("overshootFactor") is not None else 1.2,
            shader=shader, shader_analytics=shader_analytics, time=float(job.get("time", 0.0)))

    # ------------------------------------------------------------------ page-locked planes
    def _pinned_array(self, key: str, n: int, dtype) -> np.ndarray:
        nbytes = max(1, n * np.dtype(dtype).itemsize)
        ent = self._pinned.get(key)
        if ent is None or ent[1] < nbytes:
            if ent is not None:
                self._check(self._L.rm_pool_host_free(self._h, ent[0]))
            p = C.c_void_p()
            self._check(self._L.rm_pool_host_alloc(self._h, nbytes + nbytes // 8, C.byref(p)))
            ent = (p.value, nbytes + nbytes // 8)
            self._pinned[key] = ent
        buf = (C.c_uint8 * nbytes).from_address(ent[0])
        return np.frombuffer(buf, dtype=dtype, count=n)

    # ------------------------------------------------------------------ render
    def on_message(self, job: dict, shader=None, shader_analytics=None, pinned: bool = True, extras: bool = False) -> Frame:
        """One worker Job (whole frame or row band) -> its Result, rendered by all devices of the pool (rm_pool_render).
        pinned=True returns views over the pool's page-locked planes (valid until the next pinned call)."""
        rq = self._request(job, shader, shader_analytics)
        th = max(0, rq.y_end - rq.y_start)
        n = th * rq.width
        new = (lambda key, cnt, dt: self._pinned_array(key, cnt, dt)) if pinned else (lambda key, cnt, dt: np.zeros(cnt, dt))
        f = Frame(rq.y_start, rq.y_end, new("depth", n, np.uint8), new("normal", 3 * n, np.uint8), new("sdf", n, np.uint16), new("iters", n, np.uint16))
        res = _lib.Result()
        res.depth, res.normal, res.sdf_eval, res.iters = _ptr(f.depth), _ptr(f.normal), _ptr(f.sdfEval), _ptr(f.iters)
        if rq.shader >= 0:
            f.rgba = new("rgba", 4 * n, np.uint8)
            res.rgba = _ptr(f.rgba)
        if rq.shader_analytics >= 0:
            f.rgba_analytics = new("rgba2", 4 * n, np.uint8)
            res.rgba_analytics = _ptr(f.rgba_analytics)
        if extras:
            f.depth_f32, f.sdf_u32, f.depth_f64 = new("depth_f32", n, np.float32), new("sdf_u32", n, np.uint32), new("depth_f64", n, np.float64)
            res.depth_f32, res.sdf_eval_u32, res.depth_f64 = _ptr(f.depth_f32), _ptr(f.sdf_u32), _ptr(f.depth_f64)
        self._check(self._L.rm_pool_render(self._h, C.byref(rq), C.byref(res)))
        return f

    def render_device(self, job: dict, shader=None) -> dict:
        """The frame into planes in device 0's HBM; the other devices' kernels store their stripes straight into them over
        NVLink (fused gather).  Returns the frame diagnostics; `download_device_frame` copies the planes to the host."""
        rq = self._request(job, shader)
        th = max(0, rq.y_end - rq.y_start)
        n = th * rq.width
        al = lambda x: (x + 255) & ~255  # noqa: E731
        offs, o = {}, 0
        for name, size in (("depth", n), ("normal", 3 * n), ("sdf", 2 * n), ("iters", 2 * n), ("rgba", 4 * n)):
            offs[name] = o
            o += al(size)
        if self._dev_frame is None or self._dev_frame[1] < o:
            if self._dev_frame is not None:
                self._check(self._L.rm_pool_free(self._h, self._dev_frame[0]))
            p = C.c_void_p()
            self._check(self._L.rm_pool_alloc(self._h, o + 256, C.byref(p)))
            self._dev_frame = (p.value, o)
        base = self._dev_frame[0]
        res = _lib.Result()
        res.depth, res.normal, res.sdf_eval, res.iters = base + offs["depth"], base + offs["normal"], base + offs["sdf"], base + offs["iters"]
        if rq.shader >= 0:
            res.rgba = base + offs["rgba"]
        self._dev_layout = (offs, n, rq.shader >= 0)
        self._check(self._L.rm_pool_render_device(self._h, C.byref(rq), C.byref(res)))
        return self.stats()

    def download_device_frame(self) -> dict:
        offs, n, has_rgba = self._dev_layout
        out = {"depth": np.zeros(n, np.uint8), "normal": np.zeros(3 * n, np.uint8), "sdfEval": np.zeros(n, np.uint16), "iters": np.zeros(n, np.uint16)}
        if has_rgba:
            out["rgba"] = np.zeros(4 * n, np.uint8)
        for k, name in (("depth", "depth"), ("normal", "normal"), ("sdfEval", "sdf"), ("iters", "iters"), ("rgba", "rgba")):
            if k in out:
                self._check(self._L.rm_pool_memcpy_d2h(self._h, _ptr(out[k]), self._dev_frame[0] + offs[name], out[k].nbytes))
        return out

    def render_frames(self, jobs, shader=None) -> list:
        """Frame-parallel: job k on device k % n_devices, planes stay in device scratch, per-frame diagnostics come back
        (the Analytics rotation sweep, main.ts:438-441 + 527-548).  All jobs must use the scene of the first one."""
        if not jobs:
            return []
        rqs = (_lib.Request * len(jobs))()
        for k, job in enumerate(jobs):
            rqs[k] = self._request(job, shader)
        sts = (_lib.Stats * len(jobs))()
        self._check(self._L.rm_pool_render_frames(self._h, rqs, len(jobs), None, sts))
        return [_stats_dict(s) for s in sts]

    # ------------------------------------------------------------------ diagnostics
    def stats(self) -> dict:
        st = _lib.Stats()
        self._check(self._L.rm_pool_stats(self._h, C.byref(st)))
        return _stats_dict(st)

    def device_stats(self, i: int) -> dict:
        st = _lib.Stats()
        self._check(self._L.rm_pool_device_stats(self._h, i, C.byref(st)))
        return _stats_dict(st)

    def probe_fp32_peak(self) -> float:
        v = C.c_double(0.0)
        self._check(self._L.rm_pool_probe_fp32_peak(self._h, C.byref(v)))
        return v.value
