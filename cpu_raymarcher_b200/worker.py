"""RaymarchWorker — the reference's worker contract (src/workers/raymarchWorker.ts) served by the GPU.

`on_message(job) -> result` takes the same Job fields main.ts posts (main.ts:474-486) and returns the
same Result arrays (raymarchWorker.ts:24-31): depth Uint8Clamped[W*th], normal Uint8Clamped[3*W*th],
sdfEval Uint16[W*th], iters Uint16[W*th], tile-local.  Differences from the reference worker, all
invisible to the caller: the scene + acceleration structure are built once per
(preset, accelerationStructure) and kept resident in HBM instead of being rebuilt per job
(raymarchWorker.ts:37-38), and the march runs in librm_b200.so.  No CPU fallback exists: without the
library or a GPU, construction raises.
"""
from __future__ import annotations

from .renderer import Context, Frame
from .scene import Scene


class RaymarchWorker:
    def __init__(self, device: int = 0, validate_fp64: bool = False, length_sqrt: bool = False):
        self.ctx = Context(device, validate_fp64=validate_fp64, length_sqrt=length_sqrt)
        self._scene_key = None
        self.scene: Scene | None = None

    def close(self):
        self.ctx.close()

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
            # rm_upload_scene compiles operator trees and builds the BVH / octree natively
            self.ctx.upload_scene(t, m, q, accel, op_nodes=sc.primitives.op_nodes, object_root=sc.primitives.object_root)
            self.scene = sc
            self._scene_key = key
        return self.scene

    def on_message(self, job: dict, shader=None, shader_analytics=None, extras: bool = False, pinned: bool = False) -> Frame:
        """self.onmessage (raymarchWorker.ts:33-92).  `shader`/`shader_analytics` optionally fuse the
        main thread's ShadingModel.shade passes (main.ts:493-515) into the same launch."""
        width, height = int(job["width"]), int(job["height"])
        scene = self._ensure_scene(int(job.get("scenePresetIndex", 0)), job.get("accelerationStructure", "None"),
                                   job.get("synthetic"))
        cam = job.get("camera", {})
        scene.camera.set_angles(float(cam.get("pitch", 0.0)), float(cam.get("yaw", 0.0)))  # raymarchWorker.ts:39
        rq = Context.make_request(
            width, height, scene.camera.get_rotation_matrix3(), scene.camera.get_position(),
            algorithm=job.get("algorithm", "sphere-tracer"),  # unknown -> sphere tracer (raymarchWorker.ts:66-67)
            y_start=int(job.get("yStart", 0)), y_end=int(job.get("yEnd", height)),
            step_size=float(job.get("stepSize", 0.1) if job.get("stepSize") is not None else 0.1),
            overshoot=float(job.get("overshootFactor", 1.2) if job.get("overshootFactor") is not None else 1.2),
            shader=shader, shader_analytics=shader_analytics, time=float(job.get("time", 0.0)))
        return self.ctx.render(rq, extras=extras, pinned=pinned)

    def stats(self) -> dict:
        return self.ctx.stats()
