"""Row-stripe sharding of one frame over the GPUs of a box — one process per GPU (torchrun).

The reference's only parallelism is row-band data parallelism over <= 4 web workers with the tiles
copied into the full-frame buffers on the main thread and the diagnostics summed serially
(src/main.ts:318-321,444-471,527-543).  Here:

  * scene "broadcast" (each reference worker rebuilds the scene, raymarchWorker.ts:37-38): rank 0 builds
    the primitive arrays and the flattened BVH/octree once; one NCCL broadcast of the packed blob;
    every rank uploads it and keeps it resident;
  * work split: rows are dealt to ranks as interleaved 8-row stripes (a contiguous band per rank —
    the reference's rule — leaves sky rows on one GPU and dense rows on another);
  * tile "gather" (main.ts:461-468): FUSED into the render kernel.  Rank 0 owns the full-frame planes;
    every other rank maps them through CUDA IPC and its render kernel stores its pixels straight into
    rank 0's HBM over NVLink — no staging buffer, no separate gather collective;
  * diagnostics "reduce" (main.ts:527-543): each kernel's epilogue leaves 12 words per GPU; one NCCL
    all-reduce per operator combines them (sum; max, with the minima negated into the same call);
  * frames wanted in HOST memory (the main thread's frame buffers, main.ts:324-329): `render_frame_host`.
    The frame lives in one shared-memory block that every rank maps and page-locks (rm_host_register);
    each rank calls rm_render with its stripes and the shared planes, so every GPU downloads its own rows
    over its own PCIe link while its kernel is still rendering — no gather at all, N links instead of one.

With world_size == 1 every collective disappears and this is a thin wrapper over rm_render_device.
`render_band` can be replaced (tests inject a CPU renderer to exercise the partition / merge logic
under gloo without a GPU).
"""
from __future__ import annotations

import ctypes as C
import mmap
import os
import time

import numpy as np

from . import _lib
from .renderer import Context

STRIPE_ROWS = 8  # measured against 4-row stripes on the 1/8 share of cfg4: 8 rows 0.774, 4 rows 0.756 (profiles/r02k_stripe_time.log)


def stripe_rows_of(rank: int, world: int, height: int, stripe_rows: int = STRIPE_ROWS):
    """Image rows owned by `rank`: stripes rank, rank+world, ... of `stripe_rows` rows each."""
    rows = []
    n_stripes = (height + stripe_rows - 1) // stripe_rows
    for s in range(rank, n_stripes, world):
        rows.extend(range(s * stripe_rows, min((s + 1) * stripe_rows, height)))
    return np.asarray(rows, np.int32)


def plane_layout(width: int, height: int):
    """Byte offsets of the full-frame planes inside one device allocation (256-byte aligned sections)."""
    n = width * height
    al = lambda x: (x + 255) & ~255  # noqa: E731
    off, lay = 0, {}
    for name, size in (("depth", n), ("normal", 3 * n), ("sdf", 2 * n), ("iters", 2 * n), ("rgba", 4 * n)):
        lay[name] = off
        off += al(size)
    lay["total"] = off
    return lay


def reduce_stats(stats_list):
    """Combine per-rank diagnostics the way main.ts:527-543 would over the whole frame."""
    out = dict(stats_list[0])
    for s in stats_list[1:]:
        for k in ("n_pixels", "sum_sdf", "sum_iters", "sum_sdf_full", "sum_iters_full", "n_hit", "algorithmic_flops", "executed_flops", "tc_passes", "tc_requests",
                  "tc_items", "n_launches"):
            out[k] += s[k]
        for k in ("max_sdf", "max_iters"):
            out[k] = max(out[k], s[k])
        for k in ("min_sdf", "min_iters"):
            out[k] = min(out[k], s[k])
        out["evals_by_type"] = [a + b for a, b in zip(out["evals_by_type"], s["evals_by_type"])]
        out["kernel_ms"] = max(out["kernel_ms"], s["kernel_ms"])
    return out


SWEEP_SUM_KEYS = ("n_pixels", "sum_sdf", "sum_iters", "n_hit")
SWEEP_MAX_KEYS = ("max_sdf", "max_iters")
SWEEP_COLS = SWEEP_SUM_KEYS + SWEEP_MAX_KEYS + ("min_sdf", "min_iters", "_n")


def allreduce_frame_table(my_stats: dict, n_frames: int, device=None) -> list:
    """Frame-parallel sweeps (config 5): every frame is rendered whole by ONE rank, so the per-frame diagnostics of main.ts:527-548
    are complete on that rank and zero everywhere else — one SUM all-reduce of the [frames x columns] table carries every frame's
    row to every rank (the `_n` column counts the contributions: exactly 1 per frame).  `my_stats` maps frame index -> stats dict of
    the frames this rank rendered.  Returns one dict per frame, identical on every rank."""
    import torch
    import torch.distributed as dist
    table = torch.zeros((n_frames, len(SWEEP_COLS)), dtype=torch.int64)
    for k, st in my_stats.items():
        table[k] = torch.tensor([int(st[q]) for q in SWEEP_SUM_KEYS + SWEEP_MAX_KEYS] + [int(st["min_sdf"]), int(st["min_iters"]), 1])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if device is not None:
            table = table.to(device)
        dist.all_reduce(table, op=dist.ReduceOp.SUM)
        table = table.cpu()
    rows = [dict(zip(SWEEP_COLS, [int(v) for v in row])) for row in table.tolist()]
    missing = [k for k, r in enumerate(rows) if r["_n"] != 1]
    if missing:
        raise RuntimeError(f"frames {missing[:8]} were rendered by {rows[missing[0]]['_n']} ranks instead of exactly one")
    return rows


class HostFrameUnavailable(RuntimeError):
    """The shared host frame could not be created (raised on every rank alike)."""


class FrameSharder:
    def __init__(self, worker, rank: int = 0, world: int = 1, local_rank: int = 0):
        self.worker = worker
        self.ctx: Context = worker.ctx
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.frame_ptr = None      # base of the full-frame planes this rank writes into (local on rank 0, peer-mapped elsewhere)
        self.frame_owned = False
        self.layout = None
        self.size = None
        self.host_frame = None     # (mmap, uint8 view) of the shared host frame of render_frame_host
        self.host_size = None

    # ------------------------------------------------------------------ scene
    def setup_scene(self, job: dict):
        if self.world == 1:
            self.worker._ensure_scene(int(job.get("scenePresetIndex", 0)), job.get("accelerationStructure", "None"),
                                      job.get("synthetic"))
            return
        import torch
        import torch.distributed as dist
        from .renderer import OP_NODE_DTYPE, build_bvh_scene, build_octree_scene
        from .scene import Scene

        accel = job.get("accelerationStructure", "None")
        accel = accel if accel in ("Octree", "BVH") else "None"
        dev = torch.device("cpu") if dist.get_backend() == "gloo" else torch.device("cuda", self.local_rank)  # gloo: CPU tests
        # blob sections: types | world->local | params | operator-tree nodes | object roots | accel nodes | leaf lists
        if self.rank == 0:
            sc = Scene(accel)
            if job.get("synthetic"):
                sc.load_synthetic(*job["synthetic"])
            else:
                sc.load_preset(int(job.get("scenePresetIndex", 0)))
            t, m, q = sc.primitives.arrays()
            ops, roots = sc.primitives.op_nodes, sc.primitives.object_root
            has_tree = roots is not None and len(roots) > 0
            ops_b = np.ascontiguousarray(ops, OP_NODE_DTYPE).tobytes() if has_tree else b""
            roots_b = np.ascontiguousarray(roots, np.int32).tobytes() if has_tree else b""
            parts = [t.tobytes(), m.tobytes(), q.tobytes(), ops_b, roots_b, b"", b""]
            n_nodes = 0
            if accel != "None":
                # the structure indexes scene OBJECTS (operator trees included): the *_scene builders
                build = build_bvh_scene if accel == "BVH" else build_octree_scene
                nodes, n_nodes, leaf = build(t, m, q, ops if has_tree else None, roots if has_tree else None, self.ctx.flags)
                node_size = C.sizeof(_lib.BvhNode if accel == "BVH" else _lib.OctreeNode)
                parts[5] = bytes(C.string_at(C.addressof(nodes), n_nodes * node_size))
                parts[6] = np.ascontiguousarray(leaf, np.int32).tobytes()
            hdr = np.array([len(t), n_nodes] + [len(p) for p in parts], np.int64)
            blob = np.frombuffer(b"".join(parts), np.uint8)
        else:
            hdr = np.zeros(9, np.int64)
            blob = None
        hdr_t = torch.from_numpy(hdr.copy()).to(dev)
        dist.broadcast(hdr_t, src=0)
        hdr = hdr_t.cpu().numpy()
        total = int(hdr[2:].sum())
        blob_t = torch.from_numpy(blob.copy()).to(dev) if self.rank == 0 else torch.empty(total, dtype=torch.uint8, device=dev)
        dist.broadcast(blob_t, src=0)  # NCCL over NVLink: primitives + operator trees + flattened acceleration structure
        raw = blob_t.cpu().numpy().tobytes()
        self._upload_blob(raw, hdr, accel)
        self.worker.scene = Scene(accel)
        self.worker._scene_key = (int(job.get("scenePresetIndex", 0)), job.get("synthetic"), accel)

    def _upload_blob(self, raw: bytes, hdr, accel: str):
        """Every rank: unpack the broadcast scene blob and make it resident (rm_upload_scene)."""
        from .renderer import OP_NODE_DTYPE
        n, n_nodes = int(hdr[0]), int(hdr[1])
        offs = np.cumsum([0] + [int(x) for x in hdr[2:9]])
        sec = lambda i: raw[offs[i]:offs[i + 1]]  # noqa: E731
        t = np.frombuffer(sec(0), np.uint8)
        m = np.frombuffer(sec(1), np.float32).reshape(n, 16)
        q = np.frombuffer(sec(2), np.float64).reshape(n, 4)
        ops = np.frombuffer(sec(3), OP_NODE_DTYPE) if len(sec(3)) else None
        roots = np.frombuffer(sec(4), np.int32) if len(sec(4)) else None
        kw = dict(op_nodes=ops, object_root=roots)
        if accel != "None":
            node_t = _lib.BvhNode if accel == "BVH" else _lib.OctreeNode
            nodes = (node_t * max(n_nodes, 1)).from_buffer_copy(sec(5).ljust(C.sizeof(node_t), b"\0"))
            leaf = np.frombuffer(sec(6), np.int32)
            self.ctx.upload_scene(t, m, q, accel, nodes=nodes, n_nodes=n_nodes, leaf=leaf, **kw)
        else:
            self.ctx.upload_scene(t, m, q, accel, **kw)

    # ------------------------------------------------------------------ frame planes
    def _ensure_frame(self, width: int, height: int):
        if self.size == (width, height):
            return
        self.release()
        self.layout = plane_layout(width, height)
        self.size = (width, height)
        if self.world == 1:
            self.frame_ptr = self.ctx.alloc(self.layout["total"])
            self.frame_owned = True
            return
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", self.local_rank)
        if self.rank == 0:
            self.frame_ptr = self.ctx.alloc(self.layout["total"])
            self.frame_owned = True
            handle = np.frombuffer(self.ctx.ipc_export(self.frame_ptr), np.uint8).copy()
        else:
            handle = np.zeros(64, np.uint8)
        h_t = torch.from_numpy(handle).to(dev)
        dist.broadcast(h_t, src=0)
        if self.rank != 0:
            self.frame_ptr = self.ctx.ipc_open(h_t.cpu().numpy().tobytes())  # rank 0's planes, reachable over NVLink
            self.frame_owned = False

    def _release_host(self):
        if self.host_frame is not None:
            mm, view = self.host_frame
            self.ctx.host_unregister(view)
            self.host_frame = None
            self.host_size = None
            del view
            try:
                mm.close()
            except BufferError:  # a caller still holds plane views: the mapping goes when they do
                pass

    def release(self):
        self._release_host()
        if self.frame_ptr is not None:
            if self.frame_owned:
                self.ctx.free(self.frame_ptr)
            else:
                self.ctx.ipc_close(self.frame_ptr)
        self.frame_ptr = None
        self.size = None

    def _result(self, shader):
        lay, base = self.layout, self.frame_ptr
        res = _lib.Result()
        res.depth, res.normal = base + lay["depth"], base + lay["normal"]
        res.sdf_eval, res.iters = base + lay["sdf"], base + lay["iters"]
        if shader is not None:
            res.rgba = base + lay["rgba"]
        return res

    def _request(self, job: dict, shader):
        scene = self.worker.scene
        cam = job.get("camera", {})
        scene.camera.set_angles(float(cam.get("pitch", 0.0)), float(cam.get("yaw", 0.0)))
        W, H = int(job["width"]), int(job["height"])
        return Context.make_request(W, H, scene.camera.get_rotation_matrix3(), scene.camera.get_position(),
                                    algorithm=job.get("algorithm", "sphere-tracer"), y_start=0, y_end=H,
                                    step_size=float(job["stepSize"]) if job.get("stepSize") is not None else 0.1,  # only `undefined` defaults (raymarchWorker.ts)
                                    overshoot=float(job["overshootFactor"]) if job.get("overshootFactor") is not None else 1.2,
                                    shader=shader, time=float(job.get("time", 0.0)),
                                    stripes=(STRIPE_ROWS, self.world, self.rank) if self.world > 1 else None)

    # ------------------------------------------------------------------ one frame, device-resident
    def render_frame(self, job: dict, shader=None) -> dict:
        """Every rank renders its stripes straight into rank 0's planes; returns frame-level diagnostics."""
        W, H = int(job["width"]), int(job["height"])
        self._ensure_frame(W, H)
        rq = self._request(job, shader)
        self._frame_fence()
        self.ctx.render_device(rq, self._result(shader))  # synchronises this rank's stream
        return self._reduce_frame_stats(self.ctx.stats())

    def _frame_fence(self):
        """Frame k+1 must not be stored into the shared planes (rank 0's HBM, or the shared host frame) while the
        consumer still reads frame k: every rank waits here until ALL ranks — rank 0, the consumer, included — have
        come back for the next frame.  The planes a render_frame* call returns are therefore valid until the same
        rank calls render_frame* again."""
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()

    def _reduce_frame_stats(self, st: dict) -> dict:
        """Frame-level diagnostics (main.ts:527-543) from this rank's share: two NCCL all-reduces (sum, max)."""
        st["n_prims"] = self.ctx.n_prims
        if self.world == 1:
            st["frame_ms"] = st["kernel_ms"]
            st["kernel_ms_max"] = st["kernel_ms"]
            return st
        import torch
        import torch.distributed as dist
        dev = torch.device("cpu") if dist.get_backend() == "gloo" else torch.device("cuda", self.local_rank)  # gloo: CPU tests
        # two all-reduces on packed int64 words: SUM (counters; FLOP totals are integer-valued) and MAX (minima negated)
        sums = torch.tensor([st["n_pixels"], st["sum_sdf"], st["sum_iters"], st["sum_sdf_full"], st["sum_iters_full"], st["n_hit"],
                             st["n_launches"]] + list(st["evals_by_type"]) +
                            [round(st["algorithmic_flops"]), round(st["executed_flops"]), st["tc_passes"], st["tc_requests"], st["tc_items"],
                             round(st.get("fp32_pipe_flops", 0.0)), round(st.get("tensor_flops", 0.0))],
                            dtype=torch.int64, device=dev)
        maxs = torch.tensor([st["max_sdf"], st["max_iters"], int(st["kernel_ms"] * 1e6), -st["min_sdf"], -st["min_iters"]],
                            dtype=torch.int64, device=dev)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(maxs, op=dist.ReduceOp.MAX)  # also orders every rank's stores / downloads before anybody reads the frame
        both = torch.cat([sums, maxs]).tolist()  # one read-back
        s, mx = both[:17], both[17:]
        st.update(n_pixels=s[0], sum_sdf=s[1], sum_iters=s[2], sum_sdf_full=s[3], sum_iters_full=s[4], n_hit=s[5], n_launches=s[6],
                  evals_by_type=s[7:10], algorithmic_flops=float(s[10]), executed_flops=float(s[11]),
                  tc_passes=s[12], tc_requests=s[13], tc_items=s[14], fp32_pipe_flops=float(s[15]), tensor_flops=float(s[16]), max_sdf=mx[0], max_iters=mx[1], min_sdf=-mx[3], min_iters=-mx[4])
        st["kernel_ms_max"] = mx[2] / 1e6
        st["frame_ms"] = st["kernel_ms_max"]  # the gather is fused into the kernel: the slowest rank's kernel is the frame
        return st

    def download_frame(self, shader=None) -> dict:
        """Rank 0: copy the assembled full-frame planes to host arrays (the reference's frame buffers, main.ts:324-329).
        The arrays are views over page-locked memory that the next download_frame overwrites — copy what must be kept."""
        W, H = self.size
        n = W * H
        # page-locked planes owned by the context (rm_host_alloc), reused from frame to frame: the D2H runs at PCIe speed
        pa = self.ctx._pinned_array
        out = {"depth": pa("mg_depth", n, np.uint8), "normal": pa("mg_normal", 3 * n, np.uint8), "sdfEval": pa("mg_sdf", n, np.uint16),
               "iters": pa("mg_iters", n, np.uint16)}
        if shader is not None:
            out["rgba"] = pa("mg_rgba", 4 * n, np.uint8)
        for k, name in (("depth", "depth"), ("normal", "normal"), ("sdfEval", "sdf"), ("iters", "iters"), ("rgba", "rgba")):
            if k in out:
                self.ctx.memcpy_d2h(out[k], self.frame_ptr + self.layout[name])
        return out

    # ------------------------------------------------------------------ one frame into shared HOST planes
    def _ensure_host_frame(self, width: int, height: int):
        if self.host_size == (width, height):
            return
        self._release_host()
        lay = plane_layout(width, height)
        if self.world == 1:
            name = None
            mm = mmap.mmap(-1, lay["total"])
        else:
            import torch.distributed as dist
            box = [None]
            if self.rank == 0:
                try:  # a container's /dev/shm can be tiny: refuse up front rather than SIGBUS on first touch
                    vfs = os.statvfs("/dev/shm")
                    if vfs.f_bavail * vfs.f_frsize >= lay["total"] + (16 << 20):
                        box[0] = f"/dev/shm/rm_b200_frame_{os.getpid()}_{os.urandom(8).hex()}_{width}x{height}"
                        fd = os.open(box[0], os.O_CREAT | os.O_EXCL | os.O_NOFOLLOW | os.O_RDWR, 0o600)  # never an existing file / link
                        try:
                            os.ftruncate(fd, lay["total"])
                        except OSError:
                            os.close(fd)
                            os.unlink(box[0])
                            raise
                except OSError:
                    box[0] = None
            dist.broadcast_object_list(box, src=0)
            name = box[0]
            if name is None:
                raise HostFrameUnavailable(f"/dev/shm cannot hold a {lay['total'] >> 20} MiB frame")
            if self.rank != 0:
                fd = os.open(name, os.O_RDWR)
            mm = mmap.mmap(fd, lay["total"])
            os.close(fd)
            dist.barrier()  # every rank has mapped the block: the name can go, the memory stays until the last unmap
            if self.rank == 0:
                os.unlink(name)
        view = np.frombuffer(mm, np.uint8)
        self.ctx.host_register(view)
        self.host_frame = (mm, view)
        self.host_size = (width, height)
        self.host_layout = lay

    def host_planes(self, shader=None) -> dict:
        """numpy views of the shared host frame (valid until the next size change / release)."""
        W, H = self.host_size
        n, lay, v = W * H, self.host_layout, self.host_frame[1]
        out = {"depth": v[lay["depth"]:lay["depth"] + n], "normal": v[lay["normal"]:lay["normal"] + 3 * n],
               "sdfEval": v[lay["sdf"]:lay["sdf"] + 2 * n].view(np.uint16), "iters": v[lay["iters"]:lay["iters"] + 2 * n].view(np.uint16)}
        if shader is not None:
            out["rgba"] = v[lay["rgba"]:lay["rgba"] + 4 * n]
        return out

    def render_frame_host(self, job: dict, shader=None):
        """Every rank renders its stripes and downloads them itself (rm_render, band by band while the kernel runs) into
        one shared page-locked host frame.  Returns (frame diagnostics, planes); the planes are complete on every rank
        once the call returns (the stats all-reduce orders all ranks' downloads before anybody reads)."""
        W, H = int(job["width"]), int(job["height"])
        self._ensure_host_frame(W, H)
        planes = self.host_planes(shader)
        self._frame_fence()
        self.ctx.render_into(self._request(job, shader), planes)  # returns when this rank's rows are in host memory
        st = self._reduce_frame_stats(self.ctx.stats())
        return st, planes

    # ------------------------------------------------------------------ end-to-end frames (host buffers)
    def e2e_frames(self, job: dict, shader, steps: int = 3, between=None) -> dict:
        """The same frame through the reference-facing call with HOST buffers: request in, planes out.
        world == 1: RaymarchWorker.on_message (rm_render).  world > 1: render_frame_host (every rank downloads its own stripes into one shared host frame)."""
        W, H = int(job["width"]), int(job["height"])
        n = W * H
        d2h = n * (1 + 3 + 2 + 2 + (4 if shader else 0))
        h2d = C.sizeof(_lib.Request)
        if self.world == 1:
            self.worker.on_message(job, shader=shader, pinned=True)  # warm-up (allocates the page-locked planes)
            t_all = 0.0
            for _ in range(steps):
                if between:
                    between()  # e.g. an L2 flush, not timed
                t0 = time.perf_counter()
                self.worker.on_message(job, shader=shader, pinned=True)
                t_all += time.perf_counter() - t0
            ms = t_all * 1e3 / steps
            return {"ms_per_frame": ms, "h2d_bytes": h2d, "d2h_bytes": d2h,
                    "path": "RaymarchWorker.on_message -> rm_render into page-locked planes (row bands downloaded during the render)"}
        import torch.distributed as dist
        try:
            self.render_frame_host(job, shader)  # warm-up maps and page-locks the shared host frame

            def frame():
                self.render_frame_host(job, shader)  # ends with the stats all-reduces: the frame is complete on every rank
            path = "shared page-locked host frame, every rank downloads its own stripes during its render"
        except HostFrameUnavailable:
            self.render_frame(job, shader)
            if self.rank == 0:
                self.download_frame(shader)  # warm-up allocates the page-locked planes

            def frame():
                self.render_frame(job, shader)
                if self.rank == 0:
                    self.download_frame(shader)
            path = "fused peer-store gather into rank 0's HBM, then rank 0 downloads the frame"
        t_all = 0.0
        for _ in range(steps):
            if between:
                between()
            dist.barrier()
            t0 = time.perf_counter()
            frame()
            dist.barrier()  # the frame is complete on every rank
            t_all += time.perf_counter() - t0
        ms = t_all * 1e3 / steps
        return {"ms_per_frame": ms, "h2d_bytes": h2d * self.world, "d2h_bytes": d2h, "path": path}
