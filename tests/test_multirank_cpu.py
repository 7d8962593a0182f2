"""world_size-2 gloo test of the N>1 host logic (stripe partition, frame merge, shared host frame, stats reduce) with a CPU band
renderer injected in place of the GPU kernel.  The renderer here is the oracle — legitimate inside tests/ —
so the merged 2-rank frame must equal the 1-rank oracle frame bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cpu_raymarcher_b200 import multigpu
    from oracle import pyoracle as po
    W, H = 64, 44
    s = po.OracleScene().load_preset(3).build_accel("Octree").set_camera(0.1, 0.4)
    rows = multigpu.stripe_rows_of(rank, world, H)
    f = s.render_rows(W, H, rows, "sphere-tracer", nthreads=2)
    # "fused gather": every rank contributes its rows to the full frame; gloo all_reduce(SUM) of disjoint rows
    frame = {k: torch.zeros(H * W * c, dtype=torch.int32) for k, c in (("depth", 1), ("normal", 3), ("sdfEval", 1), ("iters", 1))}
    for k, c in (("depth", 1), ("normal", 3), ("sdfEval", 1), ("iters", 1)):
        src = getattr(f, k).astype(np.int32).reshape(len(rows), W * c)
        dst = frame[k].view(H, W * c)
        dst[torch.from_numpy(rows.astype(np.int64))] = torch.from_numpy(src)
        dist.all_reduce(frame[k], op=dist.ReduceOp.SUM)
    st = po.stats(f.sdfEval, f.iters)
    local = dict(n_pixels=len(rows) * W, sum_sdf=int(st["total_sdf"]), sum_iters=int(st["total_iters"]), sum_sdf_full=int(f.sdf_full.sum()),
                 sum_iters_full=int(f.iters_full.sum()), n_hit=int((f.depth_f64 < 10).sum()), algorithmic_flops=0.0, executed_flops=0.0, tc_passes=0, tc_requests=0, tc_items=0, n_launches=1,
                 max_sdf=int(st["max_sdf"]), min_sdf=int(st["min_sdf"]), max_iters=int(f.iters.max()), min_iters=int(f.iters.min()),
                 evals_by_type=[int(f.sdf_full.sum()), 0, 0], kernel_ms=1.0 + rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    # the product's own reduction (two packed all-reduces) must agree with the literal per-field merge
    import types
    sharder = multigpu.FrameSharder(types.SimpleNamespace(ctx=types.SimpleNamespace(n_prims=125)), rank, world, rank)
    red = sharder._reduce_frame_stats(dict(local))
    want = multigpu.reduce_stats(gathered)
    for k in ("n_pixels", "sum_sdf", "sum_iters", "sum_sdf_full", "sum_iters_full", "n_hit", "n_launches", "max_sdf", "min_sdf",
              "max_iters", "min_iters", "evals_by_type", "tc_passes"):
        assert red[k] == want[k], (k, red[k], want[k])
    assert red["kernel_ms_max"] == want["kernel_ms"] and red["frame_ms"] == want["kernel_ms"]
    # render_frame_host: one shared-memory frame mapped by both ranks (created by rank 0, name broadcast, unlinked once
    # mapped), each rank's renderer writes ONLY its stripes into it; the stats all-reduce orders the writes.  The GPU
    # context is replaced by a stub whose render_into is the oracle, so this is the product's host logic end to end.
    from cpu_raymarcher_b200.scene import Scene

    class StubCtx:
        n_prims = 125

        def host_register(self, arr):
            self.registered = (arr.ctypes.data, arr.nbytes)

        def host_unregister(self, arr):
            assert self.registered[0] == arr.ctypes.data

        def render_into(self, rq, planes):
            assert (rq.stripe_rows, rq.stripe_count, rq.stripe_index) == (multigpu.STRIPE_ROWS, world, rank)
            lo, hi = planes["depth"].ctypes.data, planes["rgba"].ctypes.data + planes["rgba"].nbytes if "rgba" in planes else 0
            assert self.registered[0] <= lo and (hi == 0 or hi <= self.registered[0] + self.registered[1])
            for k, c in (("depth", 1), ("normal", 3), ("sdfEval", 1), ("iters", 1)):
                planes[k].reshape(H, W * c)[rows] = getattr(f, k).reshape(len(rows), W * c)

        def stats(self):
            return dict(local)

    stub_worker = types.SimpleNamespace(ctx=StubCtx(), scene=Scene("Octree"))
    host_sharder = multigpu.FrameSharder(stub_worker, rank, world, rank)
    job = dict(width=W, height=H, camera=dict(pitch=0.1, yaw=0.4), algorithm="sphere-tracer")
    for _ in range(2):  # the second frame reuses the mapping
        st_h, planes = host_sharder.render_frame_host(job)
    assert st_h["sum_sdf"] == want["sum_sdf"] and st_h["n_pixels"] == W * H
    host = {k: np.array(v, dtype=np.int32) for k, v in planes.items()}  # every rank sees the complete frame
    for k in ("depth", "normal", "sdfEval", "iters"):
        assert np.array_equal(host[k], frame[k].numpy()), k
    del planes
    host_sharder.release()
    # setup_scene at world 2: the broadcast blob must carry the operator trees (they were dropped, so operator
    # presets rendered as a plain union on N > 1 GPUs) and the acceleration structure built over scene OBJECTS.
    from cpu_raymarcher_b200.renderer import OP_NODE_DTYPE, build_bvh_scene, build_octree_scene

    class CaptureCtx:
        flags = 0

        def upload_scene(self, t, m, q_, accel, nodes=None, n_nodes=0, leaf=None, op_nodes=None, object_root=None):
            self.got = dict(t=np.array(t), m=np.array(m), q=np.array(q_), accel=accel, n_nodes=n_nodes,
                            nodes=bytes(nodes)[: n_nodes * (40 if accel == "BVH" else 48)] if nodes is not None else b"",
                            leaf=None if leaf is None else np.array(leaf), ops=None if op_nodes is None else np.array(op_nodes),
                            roots=None if object_root is None else np.array(object_root))

    for preset, accel in ((11, "BVH"), (16, "Octree"), (17, "None"), (2, "BVH")):
        cw = types.SimpleNamespace(ctx=CaptureCtx(), scene=None, _scene_key=None)
        multigpu.FrameSharder(cw, rank, world, rank).setup_scene(dict(scenePresetIndex=preset, accelerationStructure=accel))
        got = cw.ctx.got
        sc = Scene(accel)
        sc.load_preset(preset)
        t, m, q_ = sc.primitives.arrays()
        assert np.array_equal(got["t"], t) and np.array_equal(got["m"].reshape(-1), m.reshape(-1)) and np.array_equal(got["q"].reshape(-1), q_.reshape(-1))
        ops, roots = sc.primitives.op_nodes, sc.primitives.object_root
        if roots is not None and len(roots):
            assert got["ops"].tobytes() == np.ascontiguousarray(ops, OP_NODE_DTYPE).tobytes() and np.array_equal(got["roots"], roots), preset
        else:
            assert got["ops"] is None and got["roots"] is None
        if accel != "None":
            build = build_bvh_scene if accel == "BVH" else build_octree_scene
            nodes, nn, leaf = build(t, m, q_, ops, roots, 0)
            assert got["n_nodes"] == nn and got["nodes"] == bytes(nodes)[: len(got["nodes"])] and np.array_equal(got["leaf"], leaf), (preset, accel)
        assert cw._scene_key == (preset, None, accel)
    # frame-parallel sweep (config 5): frames dealt round robin, one SUM all-reduce of the per-frame table; every rank ends up with
    # every frame's diagnostics, and a frame rendered by nobody (or twice) is an error, not a silent zero
    n_frames = 7
    mine = {k: dict(n_pixels=100 + k, sum_sdf=1000 * k + 7, sum_iters=10 * k, n_hit=k, max_sdf=50 + k, max_iters=5 + k, min_sdf=k, min_iters=1)
            for k in range(rank, n_frames, world)}
    rows = multigpu.allreduce_frame_table(mine, n_frames)
    assert len(rows) == n_frames
    for k, r in enumerate(rows):
        assert (r["n_pixels"], r["sum_sdf"], r["sum_iters"], r["n_hit"], r["max_sdf"], r["max_iters"], r["min_sdf"], r["min_iters"]) == \
            (100 + k, 1000 * k + 7, 10 * k, k, 50 + k, 5 + k, k, 1), (k, r)
    try:
        multigpu.allreduce_frame_table({k: v for k, v in mine.items() if k != 1}, n_frames)  # frame 1 rendered by nobody
        lost = False
    except RuntimeError:
        lost = True
    assert lost
    if rank == 0:
        q.put(({k: v.numpy() for k, v in frame.items()}, want))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_stripe_render_equals_single_rank():
    sys.path.insert(0, ROOT)
    from oracle import pyoracle as po
    po.lib()  # build before forking
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    frame, st = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    W, H = 64, 44
    ref = po.OracleScene().load_preset(3).build_accel("Octree").set_camera(0.1, 0.4).render(W, H, "sphere-tracer")
    assert np.array_equal(frame["depth"], ref.depth.astype(np.int32))
    assert np.array_equal(frame["normal"], ref.normal.astype(np.int32))
    assert np.array_equal(frame["sdfEval"], ref.sdfEval.astype(np.int32))
    assert np.array_equal(frame["iters"], ref.iters.astype(np.int32))
    full = po.stats(ref.sdfEval, ref.iters)
    assert st["n_pixels"] == W * H
    assert st["sum_sdf"] == full["total_sdf"] and st["sum_iters"] == full["total_iters"]
    assert st["max_sdf"] == full["max_sdf"] and st["min_sdf"] == full["min_sdf"]
    assert st["kernel_ms"] == 2.0  # max over ranks
