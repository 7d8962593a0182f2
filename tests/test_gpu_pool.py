"""rm_pool (one process drives every GPU) against the single-context path, through the C ABI (run with -m gpu).

On a box with >= 2 GPUs the pool spans the real devices; on the single-GPU test box the same code path is exercised with
two and three contexts on device 0 (rm_pool_create accepts a device more than once): separate contexts, scene replicas made by
peer copies, interleaved stripe shares, per-device downloads into one host frame, reduced diagnostics.  The bar is equality bit
for bit with the frame one context renders alone (results are partition-invariant, raymarcher.ts:73,76).
"""
import numpy as np
import pytest

from conftest import make_job

pytestmark = pytest.mark.gpu

PLANES = ("depth", "normal", "sdfEval", "iters", "rgba")
STAT_KEYS = ("n_pixels", "sum_sdf", "sum_iters", "max_sdf", "min_sdf", "max_iters", "min_iters", "sum_sdf_full", "sum_iters_full", "n_hit")


def _device_lists():
    from cpu_raymarcher_b200 import _lib
    n = _lib.lib().rm_device_count()
    lists = [[0, 0], [0, 0, 0]]
    if n >= 2:
        lists.insert(0, list(range(min(n, 4))))
    return lists


JOBS = [
    dict(W=640, H=356, preset=3, accel="Octree", alg="adaptive-step-v3"),
    dict(W=512, H=300, preset=1, accel="BVH", alg="sphere-tracer", synthetic=(3000, 0x5EED0001)),
    dict(W=333, H=97, preset=8, accel="None", alg="fixed-step"),
    dict(W=320, H=184, preset=11, accel="BVH", alg="sphere-tracer", time=1234.5),   # operator tree (SmoothSubtraction)
    dict(W=320, H=184, preset=12, accel="Octree", alg="adaptive-step-v2", time=777.0),  # animated operator tree
    dict(W=1920, H=1080, preset=2, accel="BVH", alg="sphere-tracer"),               # big enough for the early band download
]


def _job(j, **kw):
    job = make_job(j["W"], j["H"], j["preset"], j["accel"], j["alg"], 0.1, 0.5, synthetic=j.get("synthetic"), time=j.get("time", 0.0))
    job.update(kw)
    return job


@pytest.mark.parametrize("validate", [False, True])
def test_pool_frame_equals_single_context_frame(validate):
    import cpu_raymarcher_b200 as rb
    solo = rb.RaymarchWorker(device=0, validate_fp64=validate)
    for devices in _device_lists():
        pool = rb.RaymarchPool(devices, validate_fp64=validate)
        assert pool.n_devices == len(devices)
        for j in JOBS:
            job = _job(j)
            ref = solo.on_message(job, shader="phong")
            rs = solo.stats()
            for pinned in (True, False):
                got = pool.on_message(job, shader="phong", pinned=pinned)
                for k in PLANES:
                    assert np.array_equal(getattr(got, k), getattr(ref, k)), (devices, j, k, pinned)
                st = pool.stats()
                for k in STAT_KEYS:
                    assert st[k] == rs[k], (devices, j, k, st[k], rs[k])
                # one render kernel per device (+ the one-CTA tile-ordering kernel after it on frames of >= 8192 tiles per device)
                assert st["n_devices"] == len(devices) and st["n_launches"] in (len(devices), 2 * len(devices))
                shares = [pool.device_stats(i) for i in range(len(devices))]
                assert sum(s["n_pixels"] for s in shares) == j["W"] * j["H"]
                assert abs(max(s["kernel_ms"] for s in shares) - st["kernel_ms"]) < 1e-9
        pool.close()
    solo.close()


def test_pool_band_jobs_share_one_render():
    """The <= 4 row-band Jobs of a frame (main.ts:444-486): the first renders the whole frame once across the devices, the
    others are served from the pool's frame cache; every band equals the single-context band."""
    import cpu_raymarcher_b200 as rb
    solo = rb.RaymarchWorker(device=0)
    pool = rb.RaymarchPool([0, 0])
    j = JOBS[1]
    W, H = j["W"], j["H"]
    rows = -(-H // 4)  # Math.ceil(height / NUM_WORKERS)
    full = solo.on_message(_job(j), shader="iteration-heatmap")
    full_stats = solo.stats()
    for frame_no, yaw in enumerate((0.5, 0.515)):
        parts = []
        for i in range(4):
            y0, y1 = min(i * rows, H), min((i + 1) * rows, H)
            job = _job(j, yStart=y0, yEnd=y1)
            job["camera"] = dict(pitch=0.1, yaw=yaw)
            parts.append(pool.on_message(job, shader="iteration-heatmap", pinned=False))
            st = pool.stats()
            assert st["n_pixels"] == W * H, "band requests report the diagnostics of their whole frame (main.ts:527-548)"
            assert st["n_launches"] == 2, "one launch per device per FRAME, not per band"
        if frame_no == 0:
            for k in PLANES:
                assert np.array_equal(np.concatenate([getattr(p, k) for p in parts]), getattr(full, k)), k
            for k in STAT_KEYS:
                assert pool.stats()[k] == full_stats[k], k
        else:  # a new camera is a new key: re-rendered, and different from the first frame
            assert not np.array_equal(np.concatenate([p.iters for p in parts]), full.iters)
            job = _job(j)
            job["camera"] = dict(pitch=0.1, yaw=yaw)
            ref = solo.on_message(job, shader="iteration-heatmap")
            for k in PLANES:
                assert np.array_equal(np.concatenate([getattr(p, k) for p in parts]), getattr(ref, k)), k
    # a band that asks for the extension planes bypasses the cache and is still right
    job = _job(j, yStart=rows, yEnd=2 * rows)
    b = pool.on_message(job, shader="iteration-heatmap", extras=True, pinned=False)
    r = solo.on_message(job, shader="iteration-heatmap", extras=True)
    for k in PLANES + ("depth_f64", "sdf_u32"):
        assert np.array_equal(getattr(b, k), getattr(r, k)), k
    pool.close()
    solo.close()


def test_pool_device_resident_frame_fused_gather():
    """rm_pool_render_device: every device's kernel stores its stripes straight into device 0's planes."""
    import cpu_raymarcher_b200 as rb
    solo = rb.RaymarchWorker(device=0)
    for devices in _device_lists()[:2]:
        pool = rb.RaymarchPool(devices)
        for j in (JOBS[1], JOBS[5]):
            job = _job(j)
            ref = solo.on_message(job, shader="phong")
            rs = solo.stats()
            st = pool.render_device(job, shader="phong")
            got = pool.download_device_frame()
            for k in PLANES:
                assert np.array_equal(got[k], getattr(ref, k)), (devices, k)
            for k in STAT_KEYS:
                assert st[k] == rs[k], k
        pool.close()
    solo.close()


def test_pool_frame_parallel_sweep_stats():
    """rm_pool_render_frames: the Analytics rotation sweep with frames dealt to the devices; per-frame diagnostics equal the
    single-context ones frame by frame."""
    import cpu_raymarcher_b200 as rb
    from cpu_raymarcher_b200.camera import Camera
    solo = rb.RaymarchWorker(device=0)
    pool = rb.RaymarchPool(_device_lists()[0])
    cam = Camera()
    jobs = []
    for _ in range(7):
        cam.rotate_camera(0.0, 0.015)  # main.ts:438-441
        jobs.append(make_job(400, 224, 9, "None", "sphere-tracer", cam.pitch, cam.yaw))
    sts = pool.render_frames(jobs, shader="normal")
    assert len(sts) == len(jobs)
    for job, st in zip(jobs, sts):
        solo.on_message(job, shader="normal")
        rs = solo.stats()
        for k in STAT_KEYS:
            assert st[k] == rs[k], (k, st[k], rs[k])
    total = pool.stats()
    assert total["n_pixels"] == 400 * 224 * len(jobs) and total["n_devices"] == pool.n_devices
    pool.close()
    solo.close()


def test_pool_validation_build_bit_exact_vs_oracle(oracle):
    import cpu_raymarcher_b200 as rb
    from oracle import compare as cmp
    W, H = 96, 56
    pool = rb.RaymarchPool([0, 0, 0], validate_fp64=True)
    for preset, accel, alg in ((3, "Octree", "sphere-tracer"), (9, "BVH", "adaptive-step-v3"), (16, "BVH", "sphere-tracer")):
        ref = oracle.OracleScene().load_preset(preset).build_accel(accel).set_camera(0.2, 0.4).render(W, H, alg)
        f = pool.on_message(make_job(W, H, preset, accel, alg, 0.2, 0.4), extras=True, pinned=False)
        e = cmp.bit_exact(f, ref)
        assert e["all"], (preset, accel, alg, e)
    pool.close()


def test_pool_errors_are_loud():
    import cpu_raymarcher_b200 as rb
    from cpu_raymarcher_b200 import _lib
    with pytest.raises(rb.RmError) as ei:
        rb.RaymarchPool([99])
    assert ei.value.code == _lib.RM_ERR_ARG
    pool = rb.RaymarchPool([0])
    rq = rb.Context.make_request(64, 64, np.eye(3, dtype=np.float32).reshape(-1), np.zeros(3, np.float32))
    res = _lib.Result()
    buf = np.zeros(64 * 64 * 4, np.uint8)
    res.depth = res.normal = res.sdf_eval = res.iters = buf.ctypes.data
    assert pool._L.rm_pool_render(pool._h, rq, res) == _lib.RM_ERR_STATE  # before rm_pool_upload_scene
    pool.close()
