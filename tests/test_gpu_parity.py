"""GPU parity tests (run with -m gpu on a B200).  Everything goes through the C ABI (librm_b200.so).

Bars (BASELINE.json north_star):
  * fp64 validation build: hit mask, per-pixel SDF-call and iteration counters, the quantised depth/normal
    planes AND the unquantised depth (as raw doubles) are BIT-EXACT against the oracle;
  * fp32 fast path: >= 99.9 % of pixels agree on the hit mask and are within 1/255 in RGB (normal / Phong),
    depth relative error <= 1e-4 — all three together on >= 99.9 % of ALL pixels.
"""
import json
import os

import numpy as np
import pytest

from conftest import make_job

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
ALGS = ("sphere-tracer", "fixed-step", "adaptive-step", "adaptive-step-v2", "adaptive-step-v3")
RGB_TOL = 1          # 1/255
DEPTH_REL_TOL = 1e-4
PIXEL_AGREEMENT = 0.999


def _oracle_scene(oracle, preset, accel, pitch=0.0, yaw=0.0, synthetic=None):
    s = oracle.OracleScene()
    if synthetic:
        s.load_synthetic(*synthetic)
    else:
        s.load_preset(preset)
    return s.build_accel(accel).set_camera(pitch, yaw)


def assert_bit_exact(f, ref, oracle, W, H, shaders=True):
    assert np.array_equal(f.sdfEval, ref.sdfEval), "SDF-call counters differ"
    assert np.array_equal(f.iters, ref.iters), "iteration counters differ"
    assert np.array_equal(f.sdf_u32, ref.sdf_full), "un-wrapped SDF-call counters differ"
    assert np.array_equal(f.depth_f64 < 10, ref.depth_f64 < 10), "hit mask differs"
    assert np.array_equal(f.depth_f64.view(np.uint64), ref.depth_f64.view(np.uint64)), "unquantised depth differs"
    assert np.array_equal(f.depth, ref.depth), "depth bytes differ"
    assert np.array_equal(f.normal, ref.normal), "normal bytes differ"


def fast_agreement(f, ref, oracle, W, H):
    """The north-star bar, literally: fraction of ALL pixels that agree on the hit mask AND are within 1/255 in
    RGB (normal plane and Phong shade) AND have depth relative error <= 1e-4.  Also returns the depth-only
    fraction among pixels both sides call hits (diagnostic: silhouette rays whose last step hovers at EPSILON
    can stop one iteration apart in fp32)."""
    hit_ref, hit_f = ref.depth_f64 < 10, f.depth_f64 < 10
    ok = hit_ref == hit_f
    ok &= np.abs(f.normal.reshape(-1, 3).astype(int) - ref.normal.reshape(-1, 3).astype(int)).max(1) <= RGB_TOL
    if f.rgba is not None:
        want = oracle.shade("phong", ref.depth, ref.normal, ref.sdfEval, ref.iters, W, H).reshape(-1, 4)
        ok &= np.abs(f.rgba.reshape(-1, 4).astype(int) - want.astype(int)).max(1) <= RGB_TOL
    rel = np.abs(f.depth_f64 - ref.depth_f64) / np.maximum(np.abs(ref.depth_f64), 1e-12)
    ok &= rel <= DEPTH_REL_TOL
    both = hit_ref & hit_f
    depth_hits = float((rel[both] <= DEPTH_REL_TOL).mean()) if both.any() else 1.0
    return float(ok.mean()), depth_hits


# ------------------------------------------------------------------------------------------ validation build
@pytest.mark.parametrize("accel", ["None", "Octree", "BVH"])
@pytest.mark.parametrize("alg", ALGS)
@pytest.mark.parametrize("preset", [0, 2, 3, 5, 8, 9])
def test_validation_bit_exact_matrix(val_worker, oracle, preset, accel, alg):
    W, H = 96, 56
    ref = _oracle_scene(oracle, preset, accel).render(W, H, alg)
    f = val_worker.on_message(make_job(W, H, preset, accel, alg), extras=True)
    assert_bit_exact(f, ref, oracle, W, H)


@pytest.mark.parametrize("pitch,yaw", [(0.3, 0.7), (-0.6, 2.5), (1.2, 5.0), (0.0, 0.015 * 123)])
@pytest.mark.parametrize("preset,accel", [(1, "BVH"), (3, "Octree"), (4, "None"), (9, "BVH"), (5, "Octree")])
def test_validation_bit_exact_rotated_camera(val_worker, oracle, preset, accel, pitch, yaw):
    W, H = 80, 48
    ref = _oracle_scene(oracle, preset, accel, pitch, yaw).render(W, H, "sphere-tracer")
    f = val_worker.on_message(make_job(W, H, preset, accel, "sphere-tracer", pitch, yaw), extras=True)
    assert_bit_exact(f, ref, oracle, W, H)


def test_validation_axis_aligned_rays_nan_path(val_worker, oracle):
    """Centre column at yaw 0 has dir.x == 0 exactly and octree planes pass through the origin's coordinates:
    0 * Infinity = NaN flows through Math.max/min in intersectRayBox (octree.ts:200-213).  Even widths hit it."""
    W, H = 64, 64
    for preset in (2, 3):
        ref = _oracle_scene(oracle, preset, "Octree").render(W, H, "sphere-tracer")
        f = val_worker.on_message(make_job(W, H, preset, "Octree", "sphere-tracer"), extras=True)
        assert_bit_exact(f, ref, oracle, W, H)


@pytest.mark.parametrize("step", [0.01, 0.05, 0.37, 0.5])
def test_validation_fixed_step_sizes(val_worker, oracle, step):
    W, H = 64, 40
    ref = _oracle_scene(oracle, 4, "BVH").render(W, H, "fixed-step", step_size=step)
    f = val_worker.on_message(make_job(W, H, 4, "BVH", "fixed-step", step=step), extras=True)
    assert_bit_exact(f, ref, oracle, W, H)


@pytest.mark.parametrize("over", [1.0, 1.5, 2.0])
@pytest.mark.parametrize("alg", ["adaptive-step-v2", "adaptive-step-v3"])
def test_validation_overshoot_factors(val_worker, oracle, alg, over):
    W, H = 64, 40
    ref = _oracle_scene(oracle, 1, "Octree", 0.1, 0.9).render(W, H, alg, overshoot=over)
    f = val_worker.on_message(make_job(W, H, 1, "Octree", alg, 0.1, 0.9, over=over), extras=True)
    assert_bit_exact(f, ref, oracle, W, H)


@pytest.mark.parametrize("accel", ["BVH", "Octree", "None"])
def test_validation_synthetic_spheres(val_worker, oracle, accel):
    """Config-4 generator at a size the oracle finishes quickly: exercises the BVH full-scene fallback
    (scene.ts:173) and deep interval lists."""
    W, H = 64, 36
    syn = (3000, 0x5EED0001)
    ref = _oracle_scene(oracle, 1, accel, synthetic=syn).render(W, H, "sphere-tracer")
    f = val_worker.on_message(make_job(W, H, 1, accel, "sphere-tracer", synthetic=syn), extras=True)
    assert_bit_exact(f, ref, oracle, W, H)
    if accel == "BVH":
        assert (ref.sdf_full >= 3000).any(), "expected at least one full-scene fallback"


def test_validation_u16_counter_wrap(val_worker, oracle):
    syn = (70000, 0x5EED0001)
    W, H = 8, 4
    ref = _oracle_scene(oracle, 1, "None", synthetic=syn).render(W, H, "sphere-tracer")
    f = val_worker.on_message(make_job(W, H, 1, "None", "sphere-tracer", synthetic=syn), extras=True)
    assert_bit_exact(f, ref, oracle, W, H)
    assert np.all(f.sdf_u32 >= 70000) and np.array_equal(f.sdfEval, (f.sdf_u32 % 65536).astype(np.uint16))


def test_validation_mixed_rotated_primitives(oracle):
    """Boxes / tori / spheres with general rotations built by the Python host mirror (not a preset)."""
    import cpu_raymarcher_b200 as rb
    from cpu_raymarcher_b200 import scene_manager as sm
    from cpu_raymarcher_b200.camera import Camera
    pl = sm.PrimitiveList()
    sm.add_box(pl, 0.3, -0.2, 0.5, (0.4, 0.2, 0.7), rotation=(0.3, 1.1, -0.8))
    sm.add_torus(pl, -1.0, 0.4, 0.0, 0.8, rotation=(1.0, 0.0, 0.5))
    sm.add_sphere(pl, 1.0, 1.0, -1.0, 0.3)
    sm.add_box(pl, -0.6, -0.9, 0.2, (0.1, 0.5, 0.3))
    sm.add_sphere(pl, 0.0, 0.2, 1.4, 0.25, rotation=(0.1, 0.2, 0.3))
    t, m, q = pl.arrays()
    W, H = 72, 48
    cam = Camera()
    cam.set_angles(0.2, 0.6)
    for accel in ("None", "BVH", "Octree"):
        ref = oracle.OracleScene().set_prims(t, m, q).build_accel(accel).set_camera(0.2, 0.6).render(W, H, "adaptive-step-v3")
        ctx = rb.Context(0, validate_fp64=True)
        ctx.upload_scene(t, m, q, accel)
        rq = rb.Context.make_request(W, H, cam.get_rotation_matrix3(), cam.get_position(), "adaptive-step-v3")
        f = ctx.render(rq, extras=True)
        assert_bit_exact(f, ref, oracle, W, H)
        st = ctx.stats()
        assert sum(st["evals_by_type"]) == int(ref.sdf_full.sum())
        ctx.close()


def test_validation_length_sqrt_flag(oracle):
    import cpu_raymarcher_b200 as rb
    W, H = 64, 40
    oracle.lib().orc_set_length_mode(0)
    try:
        ref = _oracle_scene(oracle, 8, "BVH", 0.0, 0.8).render(W, H, "sphere-tracer")
    finally:
        oracle.lib().orc_set_length_mode(1)
    w = rb.RaymarchWorker(0, validate_fp64=True, length_sqrt=True)
    f = w.on_message(make_job(W, H, 8, "BVH", "sphere-tracer", 0.0, 0.8), extras=True)
    w.close()
    assert_bit_exact(f, ref, oracle, W, H)


def _golden_cases():
    with open(os.path.join(GOLDEN, "manifest.json")) as fh:
        return json.load(fh)["cases"]


@pytest.mark.parametrize("case", _golden_cases(), ids=lambda c: c["name"])
def test_validation_matches_committed_golden(val_worker, case):
    g = np.load(os.path.join(GOLDEN, case["name"] + ".npz"))
    job = make_job(case["W"], case["H"], case["preset"], case["accel"], case["alg"], case["pitch"], case["yaw"],
                   synthetic=tuple(case["synthetic"]) if case.get("synthetic") else None, step=case["step"], over=case["over"],
                   time=case.get("time", 0.0))
    f = val_worker.on_message(job, shader="phong", shader_analytics="sdf-heatmap", extras=True)
    for k in ("depth", "normal", "sdfEval", "iters"):
        assert np.array_equal(getattr(f, k), g[k]), k
    assert np.array_equal(f.depth_f64.view(np.uint64), g["depth_f64"].view(np.uint64))
    assert np.array_equal(f.rgba, g["phong"]) and np.array_equal(f.rgba_analytics, g["sdf_heat"])


# ------------------------------------------------------------------------------------------ worker contract
def test_band_partition_invariance_and_ragged_sizes(val_worker, oracle):
    """main.ts:444-449 partition rule with 4 and 3 workers on sizes that are not multiples of the 8x4 tile."""
    W, H = 77, 45
    ref = _oracle_scene(oracle, 3, "BVH", 0.1, 0.3).render(W, H, "sphere-tracer")
    for workers in (4, 3, 7):
        rows = -(-H // workers)
        parts = []
        for i in range(workers):
            y0, y1 = min(i * rows, H), min((i + 1) * rows, H)
            if y0 >= y1:
                continue
            f = val_worker.on_message(make_job(W, H, 3, "BVH", "sphere-tracer", 0.1, 0.3, y0, y1))
            assert (f.yStart, f.yEnd) == (y0, y1) and f.depth.size == W * (y1 - y0)
            parts.append(f)
        assert np.array_equal(np.concatenate([p.sdfEval for p in parts]), ref.sdfEval)
        assert np.array_equal(np.concatenate([p.normal for p in parts]), ref.normal)
        assert np.array_equal(np.concatenate([p.depth for p in parts]), ref.depth)
        assert np.array_equal(np.concatenate([p.iters for p in parts]), ref.iters)


def test_edge_cases_empty_band_single_pixel_unknown_algorithm(val_worker, oracle):
    f = val_worker.on_message(make_job(32, 32, 0, "None", "sphere-tracer", y0=10, y1=10))
    assert f.depth.size == 0 and f.normal.size == 0
    f = val_worker.on_message(make_job(32, 32, 0, "None", "sphere-tracer", y0=20, y1=5))  # Math.max(0, yEnd - yStart)
    assert f.depth.size == 0
    ref = _oracle_scene(oracle, 0, "None").render(1, 1)
    f = val_worker.on_message(make_job(1, 1, 0), extras=True)
    assert_bit_exact(f, ref, oracle, 1, 1)
    ref = _oracle_scene(oracle, 2, "None").render(40, 24, "sphere-tracer")
    f = val_worker.on_message(make_job(40, 24, 2, "None", "no-such-algorithm"), extras=True)  # default branch: sphere tracer
    assert_bit_exact(f, ref, oracle, 40, 24)


def test_stripe_interleave_extension(oracle):
    import cpu_raymarcher_b200 as rb
    from cpu_raymarcher_b200 import multigpu
    from cpu_raymarcher_b200 import scene_manager as sm
    from cpu_raymarcher_b200.camera import Camera
    W, H = 64, 52
    ref = _oracle_scene(oracle, 3, "Octree").render(W, H)
    t, m, q = sm.get_preset(3).arrays()
    ctx = rb.Context(0, validate_fp64=True)
    ctx.upload_scene(t, m, q, "Octree")
    cam = Camera()
    got = np.full(W * H, 0xFFFF, np.uint16)
    total = 0
    for rank in range(3):
        rq = rb.Context.make_request(W, H, cam.get_rotation_matrix3(), cam.get_position(), stripes=(8, 3, rank))
        f = ctx.render(rq)
        rows = multigpu.stripe_rows_of(rank, 3, H, 8)
        sel = (rows[:, None] * W + np.arange(W)[None, :]).ravel()
        got[sel] = f.sdfEval[sel]
        st = ctx.stats()
        assert st["n_pixels"] == len(rows) * W
        total += st["sum_sdf"]
    assert np.array_equal(got, ref.sdfEval)
    assert total == int(ref.sdfEval.astype(np.int64).sum())
    ctx.close()


def test_stats_match_main_ts_diagnostics(val_worker, oracle):
    W, H = 96, 54
    ref = _oracle_scene(oracle, 3, "BVH").render(W, H, "sphere-tracer")
    val_worker.on_message(make_job(W, H, 3, "BVH", "sphere-tracer"))
    st = val_worker.stats()
    want = oracle.stats(ref.sdfEval, ref.iters)
    assert st["n_pixels"] == W * H
    assert st["sum_sdf"] == want["total_sdf"] and st["sum_iters"] == want["total_iters"]
    assert st["max_sdf"] == want["max_sdf"] and st["min_sdf"] == want["min_sdf"]
    assert st["sum_sdf_full"] == int(ref.sdf_full.astype(np.int64).sum())
    assert st["sum_iters_full"] == int(ref.iters_full.astype(np.int64).sum())
    assert st["n_hit"] == int((ref.depth_f64 < 10).sum())
    assert st["n_launches"] == 1 and st["kernel_ms"] > 0


@pytest.mark.parametrize("shader", ["normal", "phong", "sdf-heatmap", "iteration-heatmap"])
def test_standalone_shade_entry_bit_exact(val_worker, oracle, shader):
    W, H = 64, 48
    ref = _oracle_scene(oracle, 8, "None", 0.2, 0.5).render(W, H)
    got = val_worker.ctx.shade(shader, ref.depth, ref.normal, ref.sdfEval, ref.iters, W, H)
    assert np.array_equal(got, oracle.shade(shader, ref.depth, ref.normal, ref.sdfEval, ref.iters, W, H))
    # all 2^16 counter values and all depth bytes through the heat-map / phong maps
    if shader in ("sdf-heatmap", "iteration-heatmap"):
        cnt = np.arange(65536, dtype=np.uint16)
        z8, z24 = np.zeros(65536, np.uint8), np.zeros(3 * 65536, np.uint8)
        assert np.array_equal(val_worker.ctx.shade(shader, z8, z24, cnt, cnt, 256, 256), oracle.shade(shader, z8, z24, cnt, cnt, 256, 256))


def test_phong_over_random_quantised_inputs(val_worker, fast_worker, oracle):
    rng = np.random.default_rng(7)
    n = 256 * 128
    depth = rng.integers(0, 256, n, dtype=np.uint8)
    normal = rng.integers(0, 256, 3 * n, dtype=np.uint8)
    z = np.zeros(n, np.uint16)
    want = oracle.shade("phong", depth, normal, z, z, 256, 128)
    assert np.array_equal(val_worker.ctx.shade("phong", depth, normal, z, z, 256, 128), want)
    fast = fast_worker.ctx.shade("phong", depth, normal, z, z, 256, 128)
    assert np.abs(fast.astype(int) - want.astype(int)).max() <= RGB_TOL


def test_errors_are_reported_not_swallowed():
    import cpu_raymarcher_b200 as rb
    from cpu_raymarcher_b200 import _lib
    ctx = rb.Context(0)
    rq = rb.Context.make_request(8, 8, np.eye(3, dtype=np.float32).ravel(), np.zeros(3, np.float32))
    with pytest.raises(rb.RmError) as ei:
        ctx.render(rq)
    assert ei.value.code == _lib.RM_ERR_STATE
    t = np.array([7], np.uint8)  # not sphere/box/torus: operator trees are outside the path
    with pytest.raises(rb.RmError) as ei:
        ctx.upload_scene(t, np.eye(4, dtype=np.float32).reshape(1, 16), np.zeros((1, 4)))
    assert ei.value.code == _lib.RM_ERR_UNSUPPORTED_PRIMITIVE
    ctx.upload_scene(np.array([0], np.uint8), np.eye(4, dtype=np.float32).reshape(1, 16), np.array([[1.0, 0, 0, 0]]))
    rq.y_end = 99
    with pytest.raises(rb.RmError) as ei:
        ctx.render(rq)
    assert ei.value.code == _lib.RM_ERR_ARG
    ctx.close()


# ------------------------------------------------------------------------------------------ fp32 fast path
@pytest.mark.parametrize("accel", ["None", "Octree", "BVH"])
@pytest.mark.parametrize("alg", ALGS)
@pytest.mark.parametrize("preset", [0, 1, 3, 5, 9])
def test_fast_path_tolerance_matrix(fast_worker, oracle, preset, accel, alg):
    W, H = 160, 90
    ref = _oracle_scene(oracle, preset, accel, 0.1, 0.4).render(W, H, alg)
    f = fast_worker.on_message(make_job(W, H, preset, accel, alg, 0.1, 0.4), shader="phong", extras=True)
    px, dz = fast_agreement(f, ref, oracle, W, H)
    assert px >= PIXEL_AGREEMENT, f"pixel agreement {px} (depth-only among hits {dz})"
    assert dz >= 0.99, f"depth agreement among hit pixels {dz}"


@pytest.mark.parametrize("n", [256, 257, 300, 1000, 4097, 12800, 16384, 16385, 33000, 70000])
def test_fast_path_cluster_screen_sizes(fast_worker, oracle, n):
    """Translation-only spheres behind a BVH: the all-primitives fallback runs as the tensor-core cluster screen
    (>= 256 spheres).  Sizes around the 128-sphere cluster / 128-cluster block boundaries, rotated camera."""
    W, H = 96, 54
    syn = (n, 0x5EED0001)
    ref = _oracle_scene(oracle, 1, "BVH", 0.35, 1.1, synthetic=syn).render(W, H, "sphere-tracer")
    f = fast_worker.on_message(make_job(W, H, 1, "BVH", "sphere-tracer", 0.35, 1.1, synthetic=syn), shader="phong", extras=True)
    px, dz = fast_agreement(f, ref, oracle, W, H)
    assert px >= PIXEL_AGREEMENT, f"n={n}: pixel agreement {px} (depth-only among hits {dz})"
    assert np.array_equal(f.sdfEval, ref.sdfEval) or px >= PIXEL_AGREEMENT  # counters follow the same control flow
    st = fast_worker.stats()
    if (ref.sdf_full >= n).any():  # at least one full-scene fallback happened: it went through the screen
        assert st["tc_passes"] > 0 and st["tc_requests"] > 0 and st["tc_items"] >= st["tc_requests"]
        # (that the screen executes fewer FLOPs than the brute force it replaces is asserted at the real frame size,
        #  tests/test_gpu_fullsize.py: on a 96 x 54 frame nearly every pass is an end-of-frame pass serving a handful of requests)
        assert st["executed_flops"] > 0 and st["fp32_pipe_flops"] + st["tensor_flops"] == st["executed_flops"]


@pytest.mark.parametrize("W,H,y0,y1", [(1, 1, 0, 1), (8, 4, 0, 4), (9, 5, 0, 5), (33, 17, 3, 11), (640, 3, 1, 2)])
def test_fast_path_cluster_screen_tiny_frames_and_bands(fast_worker, oracle, W, H, y0, y1):
    """Frames with fewer pixels than one CTA has lanes (idle warps must still join and leave the cooperative passes)."""
    syn = (1500, 0x5EED0001)
    ref = _oracle_scene(oracle, 1, "BVH", 0.1, 0.2, synthetic=syn).render(W, H, "sphere-tracer", y_start=y0, y_end=y1)
    f = fast_worker.on_message(make_job(W, H, 1, "BVH", "sphere-tracer", 0.1, 0.2, y0, y1, synthetic=syn), extras=True)
    assert f.depth.size == W * (y1 - y0)
    assert np.array_equal(f.sdfEval, ref.sdfEval) and np.array_equal(f.iters, ref.iters)
    assert np.array_equal(f.depth, ref.depth)


def test_fast_path_host_supplied_bvh_with_fat_leaves():
    """rm_scene.nodes: a host-built structure whose leaves hold more than two primitives (what the reference's builder only
    produces at its depth limit).  The fast point query then leaves its inline leaf records for the generic leaf loop; the
    validation build walks the same nodes literally.  Both must agree."""
    import cpu_raymarcher_b200 as rb
    from cpu_raymarcher_b200 import _lib
    from cpu_raymarcher_b200 import scene_manager as sm
    from cpu_raymarcher_b200.camera import Camera
    t, m, q = sm.synthetic_spheres(400).arrays()
    nodes, nn, leaf = rb.build_bvh(t, m, q)
    src = np.frombuffer(nodes, dtype=np.dtype([("bmin", "<f4", 3), ("bmax", "<f4", 3), ("l", "<i4"), ("r", "<i4"), ("pf", "<i4"), ("pc", "<i4")]), count=nn)
    # collapse the tree to depth 2: root -> two fat leaves (the subtrees of the root's children)
    def leaves_under(i):
        if src["l"][i] < 0 and src["r"][i] < 0:
            return list(leaf[src["pf"][i]: src["pf"][i] + src["pc"][i]])
        out = []
        for ch in (src["l"][i], src["r"][i]):
            if ch >= 0:
                out += leaves_under(ch)
        return out
    L, R = int(src["l"][0]), int(src["r"][0])
    pl, pr = leaves_under(L), leaves_under(R)
    fat = (_lib.BvhNode * 3)()
    for dst, si, first, cnt, kids in ((0, 0, 0, 0, (1, 2)), (1, L, 0, len(pl), (-1, -1)), (2, R, len(pl), len(pr), (-1, -1))):
        fat[dst].bmin[:] = [float(v) for v in src["bmin"][si]]
        fat[dst].bmax[:] = [float(v) for v in src["bmax"][si]]
        fat[dst].left, fat[dst].right = kids
        fat[dst].prim_first, fat[dst].prim_count = first, cnt
    fat_leaf = np.array(pl + pr, np.int32)
    assert len(pl) > 2 and len(pr) > 2
    W, H = 96, 54
    cam = Camera()
    cam.set_angles(0.2, 0.7)
    rq = rb.Context.make_request(W, H, cam.get_rotation_matrix3(), cam.get_position())
    frames = []
    for val in (True, False):
        ctx = rb.Context(0, validate_fp64=val)
        ctx.upload_scene(t, m, q, "BVH", nodes=fat, n_nodes=3, leaf=fat_leaf)
        frames.append(ctx.render(rq, extras=True))
        ctx.close()
    a, b = frames
    assert np.array_equal(a.sdfEval, b.sdfEval) and np.array_equal(a.iters, b.iters)
    same = (a.depth == b.depth) & (np.abs(a.normal.reshape(-1, 3).astype(int) - b.normal.reshape(-1, 3).astype(int)).max(1) <= RGB_TOL)
    assert same.mean() >= PIXEL_AGREEMENT
    assert int(a.sdf_u32.max()) >= min(len(pl), len(pr))  # the fat leaves were really evaluated


@pytest.mark.parametrize("alg", ["fixed-step", "adaptive-step-v3"])
def test_fast_path_cluster_screen_other_algorithms(fast_worker, oracle, alg):
    W, H = 80, 45
    syn = (3000, 0x5EED0001)
    ref = _oracle_scene(oracle, 1, "BVH", -0.2, 2.0, synthetic=syn).render(W, H, alg)
    f = fast_worker.on_message(make_job(W, H, 1, "BVH", alg, -0.2, 2.0, synthetic=syn), shader="phong", extras=True)
    px, dz = fast_agreement(f, ref, oracle, W, H)
    assert px >= PIXEL_AGREEMENT, f"pixel agreement {px} (depth-only among hits {dz})"


@pytest.mark.parametrize("accel", ["BVH", "Octree"])
def test_fast_path_synthetic_spheres(fast_worker, oracle, accel):
    W, H = 128, 72
    syn = (4000, 0x5EED0001)
    ref = _oracle_scene(oracle, 1, accel, synthetic=syn).render(W, H, "sphere-tracer")
    f = fast_worker.on_message(make_job(W, H, 1, accel, "sphere-tracer", synthetic=syn), shader="phong", extras=True)
    px, dz = fast_agreement(f, ref, oracle, W, H)
    assert px >= PIXEL_AGREEMENT and dz >= 0.99, (px, dz)
    # counters are not part of the fast-path bar, but they should be overwhelmingly identical
    assert (f.sdfEval == ref.sdfEval).mean() > 0.99


# ------------------------------------------------------------------------------------------ full-size properties
def test_full_size_1080p_properties(fast_worker):
    """BASELINE config 2 at its real size: size-independent properties instead of an oracle frame."""
    W, H = 1920, 1080
    job = make_job(W, H, 2, "BVH", "sphere-tracer")
    full = fast_worker.on_message(job, shader="phong", shader_analytics="iteration-heatmap", extras=True)
    st = fast_worker.stats()
    assert st["n_pixels"] == W * H
    # checksum of checksums: epilogue reduction == host reduction of the planes
    assert st["sum_sdf"] == int(full.sdfEval.astype(np.int64).sum()) and st["sum_iters"] == int(full.iters.astype(np.int64).sum())
    assert st["max_sdf"] == int(full.sdfEval.max()) and st["min_sdf"] == int(full.sdfEval.min()) == 0  # BVH misses (KAT-3)
    assert st["n_hit"] == int((full.depth_f64 < 10).sum())
    # hit pixels: 1-2 leaf primitives per query or 9 on fallback; misses terminated by the BVH have zero counters
    miss0 = full.sdf_u32 == 0
    assert np.all(full.depth[miss0] == 10) and np.all(full.iters[miss0] == 0)
    assert np.all(full.normal.reshape(-1, 3)[miss0] == 128)
    # tile-partition invariance at full size (4 bands, the reference's rule)
    rows = -(-H // 4)
    parts = [fast_worker.on_message(make_job(W, H, 2, "BVH", "sphere-tracer", y0=i * rows, y1=min((i + 1) * rows, H)),
                                    shader="phong") for i in range(4)]
    assert np.array_equal(np.concatenate([p.sdfEval for p in parts]), full.sdfEval)
    assert np.array_equal(np.concatenate([p.rgba for p in parts]), full.rgba)
    # idempotence
    again = fast_worker.on_message(job, shader="phong")
    assert np.array_equal(again.normal, full.normal) and np.array_equal(again.iters, full.iters)
    # the fused shader output is a pure function of the planes
    assert np.array_equal(fast_worker.ctx.shade("phong", full.depth, full.normal, full.sdfEval, full.iters, W, H), full.rgba)
    # left-right mirror symmetry of the scene and camera: hit mask symmetric about the centre column
    hit = (full.depth_f64 < 10).reshape(H, W)
    assert (hit[:, 1:] == hit[:, 1:][:, ::-1]).mean() > 0.9995


def test_full_size_no_accel_invariant_4k(fast_worker):
    """No-accel sphere tracer: sdfEval == N * (iters + 4 * hit) holds for every pixel at 3840x2160."""
    W, H = 3840, 2160
    f = fast_worker.on_message(make_job(W, H, 4, "None", "sphere-tracer"), extras=True)
    hit = f.depth_f64 < 10
    assert np.array_equal(f.sdf_u32, 7 * (f.iters.astype(np.uint32) + 4 * hit))
    st = fast_worker.stats()
    assert st["evals_by_type"] == [st["sum_sdf_full"], 0, 0]


def test_pinned_result_planes_match_pageable(fast_worker):
    """rm_host_alloc planes are filled by direct DMA; pageable planes go through the staging copy — same bytes."""
    job = make_job(200, 120, 3, "BVH", "adaptive-step-v3", 0.2, 0.9)
    a = fast_worker.on_message(job, shader="phong", shader_analytics="iteration-heatmap", extras=True)
    b = fast_worker.on_message(job, shader="phong", shader_analytics="iteration-heatmap", extras=True, pinned=True)
    for k in ("depth", "normal", "sdfEval", "iters", "rgba", "rgba_analytics", "depth_f32", "sdf_u32", "depth_f64"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    # a smaller frame afterwards reuses the same page-locked pool
    job2 = make_job(64, 40, 0, "None")
    c = fast_worker.on_message(job2, pinned=True)
    d = fast_worker.on_message(job2)
    assert np.array_equal(c.normal, d.normal) and c.depth.size == 64 * 40


@pytest.mark.parametrize("accel,W,H", [("BVH", 2560, 1437), ("None", 4096, 1031), ("Octree", 1920, 1080)])
def test_early_band_download_matches_pageable(fast_worker, monkeypatch, accel, W, H):
    """Big frames into page-locked planes: row bands are downloaded while the kernel still runs (the kernel flags each
    finished band). Same bytes as the pageable path and as the same call with the early download switched off."""
    job = make_job(W, H, 3, accel, "sphere-tracer")
    kw = dict(shader="phong", shader_analytics="sdf-heatmap", extras=True)
    a = fast_worker.on_message(job, **kw)
    keys = ("depth", "normal", "sdfEval", "iters", "rgba", "rgba_analytics", "depth_f32", "sdf_u32", "depth_f64")
    for rep in range(2):  # the second pass reuses the planes: stale bytes of the first must all be overwritten
        b = fast_worker.on_message(job, pinned=True, **kw)
        for k in keys:
            assert np.array_equal(getattr(a, k), getattr(b, k)), (k, rep)
        b.depth[:] = 0
        b.rgba[:] = 0
        b.depth_f64[:] = 0
    monkeypatch.setenv("RM_EARLY_COPY", "0")
    c = fast_worker.on_message(job, pinned=True, **kw)
    for k in keys:
        assert np.array_equal(getattr(a, k), getattr(c, k)), k
    # planes only (no shader, no extras): fewer planes, fewer bands
    monkeypatch.delenv("RM_EARLY_COPY")
    d = fast_worker.on_message(make_job(3840, 2160, 0, "None"), pinned=True)
    e = fast_worker.on_message(make_job(3840, 2160, 0, "None"))
    for k in ("depth", "normal", "sdfEval", "iters"):
        assert np.array_equal(getattr(d, k), getattr(e, k)), k


@pytest.mark.parametrize("memory", ["pageable", "registered"])
@pytest.mark.parametrize("W,H,world", [(64, 52, 3), (2048, 1100, 2), (1920, 1080, 8)])
def test_striped_renders_fill_one_shared_host_frame(memory, W, H, world, monkeypatch):
    """rm_render with row stripes writes ONLY the rows the request owns (include/rm.h), through every download path:
    pageable staging, direct DMA into rm_host_register'ed memory, and the early band download with strided 2-D copies
    (frames >= 16 MB).  N striped calls into one frame (what N ranks do on a shared-memory frame) == one plain call."""
    import cpu_raymarcher_b200 as rb
    from cpu_raymarcher_b200 import multigpu
    from cpu_raymarcher_b200 import scene_manager as sm
    from cpu_raymarcher_b200.camera import Camera
    t, m, q = sm.get_preset(3).arrays()
    ctx = rb.Context(0)
    ctx.upload_scene(t, m, q, "Octree")
    cam = Camera()
    mk = lambda stripes: rb.Context.make_request(W, H, cam.get_rotation_matrix3(), cam.get_position(), shader="phong", stripes=stripes)  # noqa: E731
    solo = ctx.render(mk(None))
    n = W * H
    block = np.full(12 * n, 0xAB, np.uint8)
    if memory == "registered":
        ctx.host_register(block)
    planes = {"depth": block[:n], "normal": block[n:4 * n], "sdfEval": block[4 * n:6 * n].view(np.uint16),
              "iters": block[6 * n:8 * n].view(np.uint16), "rgba": block[8 * n:12 * n]}
    for early in ("1", "0"):
        monkeypatch.setenv("RM_EARLY_COPY", early)
        block[:] = 0xAB
        total = 0
        for rank in range(world):
            ctx.render_into(mk((8, world, rank)), planes)
            total += ctx.stats()["sum_sdf"]
            if rank == 0:  # nobody else's rows were touched
                other = np.setdiff1d(np.arange(H), multigpu.stripe_rows_of(0, world, H, 8))
                assert (planes["rgba"].reshape(H, -1)[other] == 0xAB).all() and (planes["sdfEval"].reshape(H, -1)[other] == 0xABAB).all()
                own = multigpu.stripe_rows_of(0, world, H, 8)
                assert np.array_equal(planes["normal"].reshape(H, -1)[own], solo.normal.reshape(H, -1)[own])
        for k in ("depth", "normal", "sdfEval", "iters", "rgba"):
            assert np.array_equal(planes[k], getattr(solo, k)), (k, early)
        assert total == int(solo.sdfEval.astype(np.int64).sum())
    if memory == "registered":
        ctx.host_unregister(block)
    ctx.close()


def test_multi_gpu_fused_gather_matches_single_gpu():
    """N >= 2 GPUs only: torchrun, one process per GPU, CUDA-IPC fused gather + NCCL stats all-reduce."""
    import subprocess
    import sys
    from cpu_raymarcher_b200 import _lib
    n = _lib.lib().rm_device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(root, "tools", "multigpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MISMATCH" not in r.stdout


# ------------------------------------------------------------------------------------------ random primitive lists
def _random_scene(rng, n):
    from cpu_raymarcher_b200 import scene_manager as sm
    pl = sm.PrimitiveList()
    for _ in range(n):
        pos = rng.uniform(-2.0, 2.0, 3)
        rot = tuple(rng.uniform(-3.0, 3.0, 3)) if rng.random() < 0.6 else None
        k = rng.integers(0, 3)
        if k == 0:
            sm.add_sphere(pl, *pos, rng.uniform(0.05, 0.5), rot)
        elif k == 1:
            sm.add_box(pl, *pos, tuple(rng.uniform(0.05, 0.4, 3)), rot)
        else:
            sm.add_torus(pl, *pos, rng.uniform(0.1, 0.5), rot)
        if rng.random() < 0.3:  # non-uniform scale folded into world->local, like createMandelbulb's mat4.scale (sceneManager.ts:62-63)
            from cpu_raymarcher_b200 import glmatrix as gm
            pl.world_to_local[-1] = gm.mat4_scale(pl.world_to_local[-1], tuple(rng.uniform(0.5, 2.0, 3)))
    return pl.arrays()


@pytest.mark.parametrize("seed", range(24))
def test_validation_random_primitive_lists(oracle, seed):
    """Fuzz: random mixes of rotated spheres / boxes / tori, every structure and algorithm, random camera and step sizes."""
    import cpu_raymarcher_b200 as rb
    from cpu_raymarcher_b200.camera import Camera
    rng = np.random.default_rng(7000 + seed)
    t, m, q = _random_scene(rng, int(rng.integers(1, 80)))
    accel = ["None", "Octree", "BVH"][seed % 3]
    alg = ALGS[(seed // 3) % len(ALGS)]
    pitch, yaw = float(rng.uniform(-1.2, 1.2)), float(rng.uniform(0, 6.28))
    step, over = float(rng.uniform(0.02, 0.3)), float(rng.uniform(1.0, 2.0))
    W, H = 72, 40
    ref = oracle.OracleScene().set_prims(t, m, q).build_accel(accel).set_camera(pitch, yaw).render(W, H, alg, step_size=step, overshoot=over)
    ctx = rb.Context(0, validate_fp64=True)
    ctx.upload_scene(t, m, q, accel)
    cam = Camera()
    cam.set_angles(pitch, yaw)
    f = ctx.render(rb.Context.make_request(W, H, cam.get_rotation_matrix3(), cam.get_position(), algorithm=alg, step_size=step, overshoot=over),
                   extras=True)
    ctx.close()
    assert_bit_exact(f, ref, oracle, W, H)


@pytest.mark.parametrize("seed", range(12))
def test_fast_path_random_primitive_lists(oracle, seed):
    """Same generator through the fp32 fast path; sizes 300-700 behind a BVH go through the CTA-cooperative all-primitives
    pass of the GENERAL record layout (8-warp CTAs, FFMA search)."""
    import cpu_raymarcher_b200 as rb
    from cpu_raymarcher_b200.camera import Camera
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.integers(300, 700)) if seed % 2 == 0 else int(rng.integers(1, 120))
    t, m, q = _random_scene(rng, n)
    accel = ["BVH", "Octree", "None"][seed % 3] if seed % 2 else "BVH"
    alg = ALGS[seed % len(ALGS)]
    pitch, yaw = float(rng.uniform(-1.0, 1.0)), float(rng.uniform(0, 6.28))
    W, H = 96, 54
    ref = oracle.OracleScene().set_prims(t, m, q).build_accel(accel).set_camera(pitch, yaw).render(W, H, alg)
    ctx = rb.Context(0)
    ctx.upload_scene(t, m, q, accel)
    cam = Camera()
    cam.set_angles(pitch, yaw)
    f = ctx.render(rb.Context.make_request(W, H, cam.get_rotation_matrix3(), cam.get_position(), algorithm=alg, shader="phong"), extras=True)
    ctx.close()
    px, dz = fast_agreement(f, ref, oracle, W, H)
    assert px >= PIXEL_AGREEMENT, f"n={n} {accel} {alg}: pixel agreement {px} (depth-only among hits {dz})"


@pytest.mark.parametrize("seed", range(16))
def test_fast_path_cluster_screen_adversarial_sphere_sets(oracle, seed):
    """Translation-only sphere sets that stress the bounds of the cluster screen: tight clumps, nested and coincident
    spheres, radii spanning three orders of magnitude, far outliers, a camera inside the cloud."""
    import cpu_raymarcher_b200 as rb
    from cpu_raymarcher_b200 import scene_manager as sm
    from cpu_raymarcher_b200.camera import Camera
    rng = np.random.default_rng(4000 + seed)
    n = int(rng.integers(256, 2500))
    kind = seed % 4
    if kind == 0:    # a few tight clumps
        centres = rng.normal(0, 0.05, (n, 3)) + rng.uniform(-1.5, 1.5, (8, 3))[rng.integers(0, 8, n)]
        radii = rng.uniform(0.005, 0.05, n)
    elif kind == 1:  # nested + coincident spheres, radii over three orders of magnitude
        centres = rng.uniform(-1.0, 1.0, (n, 3))
        centres[: n // 4] = centres[n // 4: 2 * (n // 4)]
        radii = 10 ** rng.uniform(-3, -0.3, n)
    elif kind == 2:  # a thin sheet plus far outliers
        centres = np.stack([rng.uniform(-2, 2, n), rng.normal(0, 0.01, n), rng.uniform(-2, 2, n)], 1)
        centres[:5] = rng.uniform(6, 9, (5, 3))
        radii = rng.uniform(0.01, 0.08, n)
    else:            # uniform cloud, big radii variance
        centres = rng.uniform(-2.5, 2.5, (n, 3))
        radii = np.where(rng.random(n) < 0.02, rng.uniform(0.3, 0.8, n), rng.uniform(0.01, 0.04, n))
    w2l = sm.get_transform_batch(centres)
    t = np.zeros(n, np.uint8)
    q = np.zeros((n, 4))
    q[:, 0] = radii
    pitch, yaw = float(rng.uniform(-0.6, 0.6)), float(rng.uniform(0, 6.28))
    W, H = 80, 45
    ref = oracle.OracleScene().set_prims(t, w2l, q).build_accel("BVH").set_camera(pitch, yaw).render(W, H, "sphere-tracer")
    ctx = rb.Context(0)
    ctx.upload_scene(t, w2l, q, "BVH")
    cam = Camera()
    cam.set_angles(pitch, yaw)
    f = ctx.render(rb.Context.make_request(W, H, cam.get_rotation_matrix3(), cam.get_position(), shader="phong"), extras=True)
    st = ctx.stats()
    ctx.close()
    px, dz = fast_agreement(f, ref, oracle, W, H)
    assert px >= PIXEL_AGREEMENT, f"kind {kind} n={n}: pixel agreement {px} (depth-only among hits {dz}); tc passes {st['tc_passes']}"
    assert np.array_equal(f.sdfEval, ref.sdfEval) or px >= PIXEL_AGREEMENT


@pytest.mark.parametrize("cap", [1, 3, 8])
def test_fast_path_lazy_walk_hand_over_to_literal_interval_list(oracle, cap, monkeypatch):
    """The fast BVH kernels produce the reference's sorted interval list lazily through a bounded buffer; when it overflows the
    ray hands over to the literal findRayIntersections list at the same cursor index.  RM_LAZY_CAP (read at upload) shrinks
    the buffer so that the hand-over happens on nearly every ray.  (A stale-buffer bug in this path was found with it.)"""
    import cpu_raymarcher_b200 as rb
    from cpu_raymarcher_b200 import scene_manager as sm
    from cpu_raymarcher_b200.camera import Camera
    monkeypatch.setenv("RM_LAZY_CAP", str(cap))
    W, H = 96, 54
    for syn, preset in (((1500, 0x5EED0001), 1), (None, 3)):
        t, m, q = (sm.synthetic_spheres(*syn) if syn else sm.get_preset(preset)).arrays()
        ref = _oracle_scene(oracle, preset, "BVH", 0.2, 0.9, synthetic=syn).render(W, H, "sphere-tracer")
        ctx = rb.Context(0)
        ctx.upload_scene(t, m, q, "BVH")
        cam = Camera()
        cam.set_angles(0.2, 0.9)
        f = ctx.render(rb.Context.make_request(W, H, cam.get_rotation_matrix3(), cam.get_position(), shader="phong"), extras=True)
        ctx.close()
        assert np.array_equal(f.sdfEval, ref.sdfEval) and np.array_equal(f.iters, ref.iters), (cap, preset)
        px, dz = fast_agreement(f, ref, oracle, W, H)
        assert px >= PIXEL_AGREEMENT


def test_fast_path_cluster_screen_more_blocks_than_ring_stages(fast_worker, oracle):
    """150 000 spheres = 1172 clusters = 10 tensor-core blocks: more than the 8 B-tile stages, so the ring is refilled while
    the sweeps run (config 4 itself has 7 blocks and never refills)."""
    W, H = 48, 27
    syn = (150000, 0x5EED0001)
    ref = _oracle_scene(oracle, 1, "BVH", 0.3, 0.8, synthetic=syn).render(W, H, "sphere-tracer")
    f = fast_worker.on_message(make_job(W, H, 1, "BVH", "sphere-tracer", 0.3, 0.8, synthetic=syn), shader="phong", extras=True)
    px, dz = fast_agreement(f, ref, oracle, W, H)
    assert px >= PIXEL_AGREEMENT, f"pixel agreement {px} (depth-only among hits {dz})"
    assert np.array_equal(f.sdfEval, ref.sdfEval)
    assert fast_worker.stats()["tc_passes"] > 0


def test_fast_path_cluster_screen_item_list_overflow(monkeypatch):
    """Two clumps of 60 000 spheres with a common centre each: every cluster of the nearer clump is within reach of every
    query, so a 128-request pass wants ~60 000 work items — more than the shared-memory list holds; the pass then falls back
    to evaluating every sphere for its requests.  Checked against the other all-primitives implementation of the library
    (RM_DISABLE_TC=1: the FFMA screened search), which shares no code with the cluster screen."""
    import cpu_raymarcher_b200 as rb
    from cpu_raymarcher_b200 import scene_manager as sm
    from cpu_raymarcher_b200.camera import Camera
    rng = np.random.default_rng(77)
    n = 120000
    centres = np.zeros((n, 3))
    centres[: n // 2, 0] = -1.5
    centres[n // 2:, 0] = 1.5
    centres += rng.normal(0, 1e-4, (n, 3))
    radii = rng.uniform(0.01, 0.3, n)
    w2l = sm.get_transform_batch(centres)
    t = np.zeros(n, np.uint8)
    q = np.zeros((n, 4))
    q[:, 0] = radii
    W, H = 256, 144
    cam = Camera()
    cam.set_angles(0.1, 0.3)
    rq = rb.Context.make_request(W, H, cam.get_rotation_matrix3(), cam.get_position(), shader="phong")
    frames, stats = [], []
    for disable in ("", "1"):
        if disable:
            monkeypatch.setenv("RM_DISABLE_TC", "1")
        ctx = rb.Context(0)
        ctx.upload_scene(t, w2l, q, "BVH")
        frames.append(ctx.render(rq, extras=True))
        stats.append(ctx.stats())
        ctx.close()
    a, b = frames
    assert stats[0]["tc_passes"] > 0 and stats[1]["tc_passes"] == 0
    assert stats[0]["tc_items"] / stats[0]["tc_passes"] > 3000, (stats[0]["tc_passes"], stats[0]["tc_items"])  # average; full passes hit the 7 168-item cap (RM_TC_ITEM_CAP) and fall back
    assert np.array_equal(a.sdfEval, b.sdfEval) and np.array_equal(a.iters, b.iters)
    assert np.array_equal(a.depth, b.depth) and np.array_equal(a.normal, b.normal) and np.array_equal(a.rgba, b.rgba)


def test_full_size_cluster_screen_equals_ffma_search(monkeypatch):
    """BASELINE config 4 at its full size (100 000 spheres, 3840 x 2160), where the oracle is far too slow: the tensor-core
    cluster screen and the FFMA screened search (RM_DISABLE_TC=1) are independent implementations of the all-primitives
    fallback — every plane of the frame and every diagnostic must be identical."""
    import cpu_raymarcher_b200 as rb
    from cpu_raymarcher_b200 import scene_manager as sm
    from cpu_raymarcher_b200.camera import Camera
    t, m, q = sm.synthetic_spheres(100000).arrays()
    W, H = 3840, 2160
    cam = Camera()
    rq = rb.Context.make_request(W, H, cam.get_rotation_matrix3(), cam.get_position(), shader="iteration-heatmap")
    frames, stats = [], []
    for disable in ("", "1"):
        if disable:
            monkeypatch.setenv("RM_DISABLE_TC", "1")
        ctx = rb.Context(0)
        ctx.upload_scene(t, m, q, "BVH")
        frames.append(ctx.render(rq))
        stats.append(ctx.stats())
        ctx.close()
    a, b = frames
    assert stats[0]["tc_passes"] > 0 and stats[1]["tc_passes"] == 0
    for k in ("depth", "normal", "sdfEval", "iters", "rgba"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    for k in ("sum_sdf", "sum_iters", "max_sdf", "min_sdf", "sum_sdf_full", "n_hit"):
        assert stats[0][k] == stats[1][k], k


def test_full_size_lazy_grid_walk_equals_literal_interval_list(monkeypatch):
    """Config 4 scene at 1920 x 1080: the lazily produced, grid-ordered interval list against the literal
    BVH.findRayIntersections list (RM_LAZY_CAP=1 makes every ray hand over to it): identical frames."""
    import cpu_raymarcher_b200 as rb
    from cpu_raymarcher_b200 import scene_manager as sm
    from cpu_raymarcher_b200.camera import Camera
    t, m, q = sm.synthetic_spheres(100000).arrays()
    W, H = 1920, 1080
    cam = Camera()
    cam.set_angles(0.2, 1.0)
    rq = rb.Context.make_request(W, H, cam.get_rotation_matrix3(), cam.get_position(), shader="phong")
    frames = []
    for cap in ("", "1"):
        if cap:
            monkeypatch.setenv("RM_LAZY_CAP", cap)
        ctx = rb.Context(0)
        ctx.upload_scene(t, m, q, "BVH")
        frames.append(ctx.render(rq))
        ctx.close()
    a, b = frames
    for k in ("depth", "normal", "sdfEval", "iters", "rgba"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k


@pytest.mark.parametrize("accel", ["None", "Octree", "BVH"])
def test_empty_scene(oracle, accel):
    """Scene.objectSDFs = []: getDistance returns MAX_DIST every time (scene.ts:145-146,183-189); both builds must march the
    same way as the reference (no primitive arrays, empty acceleration structures)."""
    import cpu_raymarcher_b200 as rb
    from cpu_raymarcher_b200.camera import Camera
    t, m, q = np.zeros(0, np.uint8), np.zeros((0, 16), np.float32), np.zeros((0, 4))
    W, H = 40, 24
    cam = Camera()
    for alg in ("sphere-tracer", "fixed-step", "adaptive-step-v3"):
        ref = oracle.OracleScene().set_prims(t, m, q).build_accel(accel).set_camera(0.0, 0.0).render(W, H, alg)
        for val in (True, False):
            ctx = rb.Context(0, validate_fp64=val)
            ctx.upload_scene(t, m, q, accel)
            f = ctx.render(rb.Context.make_request(W, H, cam.get_rotation_matrix3(), cam.get_position(), algorithm=alg), extras=True)
            ctx.close()
            assert np.array_equal(f.sdfEval, ref.sdfEval) and np.array_equal(f.iters, ref.iters), (accel, alg, val)
            assert np.array_equal(f.depth, ref.depth) and np.array_equal(f.normal, ref.normal), (accel, alg, val)


def test_cost_ordered_tile_queue_does_not_change_the_frame(monkeypatch):
    """From the second frame of a geometry on, tiles are handed out most-expensive-first by the previous frame's per-tile cost
    (rm_api.cu, order_tiles_kernel).  The image must not depend on the schedule: frames 1-3 (identity order, then ordered by
    frame 1's and frame 2's costs), a frame with a moved camera rendered under the stale order, and RM_TILE_ORDER=0 are all
    compared plane by plane; the order array itself must be a permutation (every pixel written exactly once is implied by the
    equality of every plane and of the diagnostics)."""
    import cpu_raymarcher_b200 as rb
    from cpu_raymarcher_b200 import scene_manager as sm
    from cpu_raymarcher_b200.camera import Camera
    t, m, q = sm.synthetic_spheres(10000).arrays()
    W, H = 1920, 1080
    cam = Camera()
    monkeypatch.setenv("RM_ANATOMY", "1")
    # (the order is used for striped requests — the shares of a multi-GPU frame; here: GPU 0 of 2)
    rq = rb.Context.make_request(W, H, cam.get_rotation_matrix3(), cam.get_position(), shader="iteration-heatmap", stripes=(8, 2, 0))
    cam.set_angles(0.1, 0.3)
    rq2 = rb.Context.make_request(W, H, cam.get_rotation_matrix3(), cam.get_position(), shader="iteration-heatmap", stripes=(8, 2, 0))
    ctx = rb.Context(0)
    ctx.upload_scene(t, m, q, "BVH")
    frames = [ctx.render(rq) for _ in range(3)]
    st = ctx.stats()
    assert st["n_launches"] == 2  # render kernel + the one-CTA ordering kernel
    moved = ctx.render(rq2)  # ordered by the costs of the OTHER camera's frame
    ctx.close()
    monkeypatch.setenv("RM_TILE_ORDER", "0")
    ctx = rb.Context(0)
    ctx.upload_scene(t, m, q, "BVH")
    plain, plain_moved = ctx.render(rq), ctx.render(rq2)
    assert ctx.stats()["n_launches"] == 1
    ctx.close()
    for k in ("depth", "normal", "sdfEval", "iters", "rgba"):
        for f in frames:
            assert np.array_equal(getattr(f, k), getattr(plain, k)), k
        assert np.array_equal(getattr(moved, k), getattr(plain_moved, k)), k
    assert int(plain.iters.astype(np.int64).sum()) > 0
    # and the frame-anatomy timers of rm_stats are filled
    assert st["drain_ms"] > 0 and st["tail_ms"] >= 0 and st["drain_ms"] + st["tail_ms"] <= st["kernel_ms"] * 1.05 + 0.05


@pytest.mark.parametrize("trigger", ["1", "4", "64"])
def test_cooperative_pass_rendezvous_stress(monkeypatch, oracle, trigger):
    """The CTA rendezvous of the all-primitives pass (request ring, go / stuck / finished flags, tensor-core pass with all 16 warps)
    under the shapes that stress it: many tiny and odd-sized frames (fewer tiles than warps, one-pixel-wide and one-row bands,
    frames that are all end-of-frame), every end-of-frame trigger setting, repeated back to back on one context.  compute-sanitizer
    is closed on this pool, so this is the protocol's safety net: every frame must equal the frame of the plain setting, and the
    first of each size the oracle."""
    import cpu_raymarcher_b200 as rb
    monkeypatch.setenv("RM_TAIL_TRIGGER", trigger)
    w = rb.RaymarchWorker(device=0)
    syn = (3000, 0x5EED0001)
    sizes = [(8, 4), (1, 1), (33, 7), (130, 61), (257, 3), (3, 129), (64, 64), (517, 11), (96, 54)]
    for k, (W, H) in enumerate(sizes):
        ref = _oracle_scene(oracle, 1, "BVH", 0.2, 0.3 * k, synthetic=syn).render(W, H, "sphere-tracer")
        first = None
        for rep in range(6):
            f = w.on_message(make_job(W, H, 1, "BVH", "sphere-tracer", 0.2, 0.3 * k, synthetic=syn), shader="phong", extras=True)
            if first is None:
                first = f
                px, dz = fast_agreement(f, ref, oracle, W, H)
                assert px >= PIXEL_AGREEMENT or W * H < 2000, (W, H, px)
                assert np.array_equal(f.sdfEval, ref.sdfEval) or px >= PIXEL_AGREEMENT
            else:
                for name in ("depth", "normal", "sdfEval", "iters", "rgba"):
                    assert np.array_equal(getattr(f, name), getattr(first, name)), (W, H, rep, name)
        # bands of one and two rows through the same context (Job.yStart / yEnd), against the full frame
        if H >= 3:
            b = w.on_message(make_job(W, H, 1, "BVH", "sphere-tracer", 0.2, 0.3 * k, y0=1, y1=3, synthetic=syn), shader="phong")
            assert np.array_equal(b.iters, first.iters[W:3 * W]) and np.array_equal(b.rgba, first.rgba[4 * W:12 * W])
    w.close()
