"""GPU parity tests of the SDF-operator widening, run with -m gpu on a B200, through the
C ABI.  Operator trees (Round, Twist, SmoothUnion, SmoothSubtraction, Repetition, AnimatedTranslate; presets 6,
10-12, 14-18) run in JS-number arithmetic in BOTH builds, so the bar is bit-exactness against the oracle for the
validation build and the north-star tolerance for the default build (which only swaps V8's compensated hypot
for a plain sqrt).  Math.sin / Math.cos of the Twist operator are libm calls on both sides (CUDA vs glibc, each
<= 1-2 ulp); their result is consumed through a float32 store, so a last-bit difference would surface with
probability ~2^-29 per evaluation — the frames below are deterministic and bit-identical."""
import numpy as np
import pytest

import cpu_raymarcher_b200 as rb
from cpu_raymarcher_b200 import _lib
from cpu_raymarcher_b200 import scene_manager as sm
from cpu_raymarcher_b200.camera import Camera

from conftest import make_job
from test_gpu_parity import ALGS, PIXEL_AGREEMENT, assert_bit_exact, fast_agreement

pytestmark = pytest.mark.gpu

OPS = sm.OPERATOR_PRESETS


def _oracle(oracle, preset, accel, pitch=0.0, yaw=0.0, time=0.0):
    return oracle.OracleScene().load_preset(preset).build_accel(accel).set_camera(pitch, yaw).set_time(time)


@pytest.mark.parametrize("accel", ["None", "Octree", "BVH"])
@pytest.mark.parametrize("preset", OPS)
def test_validation_operator_presets_bit_exact(val_worker, oracle, preset, accel):
    W, H = 96, 56
    ref = _oracle(oracle, preset, accel).render(W, H, "sphere-tracer")
    f = val_worker.on_message(make_job(W, H, preset, accel, "sphere-tracer"), extras=True)
    assert_bit_exact(f, ref, oracle, W, H)
    assert (ref.depth_f64 < 10).any(), "the preset should be visible"


@pytest.mark.parametrize("alg", ALGS)
@pytest.mark.parametrize("preset,accel", [(11, "BVH"), (16, "None"), (17, "Octree"), (18, "BVH")])
def test_validation_operator_presets_all_algorithms(val_worker, oracle, preset, accel, alg):
    W, H = 80, 48
    ref = _oracle(oracle, preset, accel, 0.25, 0.8).render(W, H, alg)
    f = val_worker.on_message(make_job(W, H, preset, accel, alg, 0.25, 0.8), extras=True)
    assert_bit_exact(f, ref, oracle, W, H)


@pytest.mark.parametrize("time", [0.0, 100.0, 314.0, 1234.5])
def test_animated_translate_follows_job_time(val_worker, oracle, time):
    """Job.time -> Scene.updateTime (raymarcher.ts:59) -> AnimatedTranslate.setTime (preset 12)."""
    W, H = 96, 56
    ref = _oracle(oracle, 12, "BVH", time=time).render(W, H, "sphere-tracer")
    f = val_worker.on_message(make_job(W, H, 12, "BVH", "sphere-tracer", time=time), extras=True)
    assert_bit_exact(f, ref, oracle, W, H)


def test_animation_changes_the_frame(val_worker):
    a = val_worker.on_message(make_job(64, 40, 12, "None", time=0.0))
    b = val_worker.on_message(make_job(64, 40, 12, "None", time=250.0))
    assert not np.array_equal(a.depth, b.depth)


def _custom_objects():
    """Operators nested in ways the presets do not: rotated leaves under Round / Twist, a repetition of a rounded
    box, an animated subtraction, next to a plain primitive."""
    rounded = sm.create_round(sm.create_box(0.3, -0.2, 0.1, (0.3, 0.2, 0.25), (0.4, 0.2, -0.3)), 0.07)
    twisted = sm.create_twist(sm.create_torus(-0.9, 0.5, 0.0, 0.5, (0.3, 0.0, 0.9)), 2.5)
    blob = sm.create_smooth_union(rounded, twisted, 0.15)
    moving = sm.create_animated_translate(sm.create_sphere(0.2, 0.9, 0.3, 0.35), (0.3, 1.0, -0.2), 0.6, 0.01)
    carved = sm.create_smooth_subtract(sm.create_box(1.2, 0.1, -0.3, (0.5, 0.5, 0.5)), moving, 0.05)
    lattice = sm.create_repetition(sm.create_round(sm.create_box(0, 0, 0, (0.08, 0.08, 0.08)), 0.02), (2.5, 2.5, 2.5))
    return [blob, carved, sm.create_sphere(-0.3, -1.1, 0.4, 0.3), lattice]


@pytest.mark.parametrize("accel", ["None", "Octree", "BVH"])
def test_validation_custom_nested_trees_bit_exact(oracle, accel):
    W, H = 96, 56
    pl = sm.flatten(_custom_objects())
    t, m, q = pl.arrays()
    osc = oracle.OracleScene().set_tree(t, m, q, pl.op_nodes, pl.object_root).build_accel(accel).set_camera(0.2, 0.6).set_time(77.0)
    ref = osc.render(W, H, "adaptive-step-v3")
    ctx = rb.Context(0, validate_fp64=True)
    ctx.upload_scene(t, m, q, accel, op_nodes=pl.op_nodes, object_root=pl.object_root)
    cam = Camera()
    cam.set_angles(0.2, 0.6)
    rq = rb.Context.make_request(W, H, cam.get_rotation_matrix3(), cam.get_position(), algorithm="adaptive-step-v3", time=77.0,
                                 shader="phong", shader_analytics="iteration-heatmap")
    f = ctx.render(rq, extras=True)
    assert_bit_exact(f, ref, oracle, W, H)
    assert np.array_equal(f.rgba, oracle.shade("phong", ref.depth, ref.normal, ref.sdfEval, ref.iters, W, H))
    assert np.array_equal(f.rgba_analytics, oracle.shade("iteration-heatmap", ref.depth, ref.normal, ref.sdfEval, ref.iters, W, H))
    ctx.close()


@pytest.mark.parametrize("preset", OPS)
def test_default_build_operator_presets_tolerance(fast_worker, oracle, preset):
    """Contexts without RM_F_VALIDATE_FP64 run operator trees in the same fp64 arithmetic (plain sqrt for
    vec3.length): the north-star bar, over all pixels."""
    W, H = 160, 90
    ref = _oracle(oracle, preset, "BVH", 0.1, 0.4).render(W, H, "sphere-tracer")
    f = fast_worker.on_message(make_job(W, H, preset, "BVH", "sphere-tracer", 0.1, 0.4), shader="phong", extras=True)
    px, dz = fast_agreement(f, ref, oracle, W, H)
    assert px >= PIXEL_AGREEMENT, f"pixel agreement {px} (depth-only among hits {dz})"


def test_operator_stats_and_flop_accounting(val_worker, oracle):
    """Preset 18 ("67"): two objects, leaves box+torus and box+box.  Every object evaluation is one SDF call
    (main.ts:527-548 counts those); evals_by_type counts the leaves underneath."""
    W, H = 64, 40
    ref = _oracle(oracle, 18, "None").render(W, H, "sphere-tracer")
    val_worker.on_message(make_job(W, H, 18, "None", "sphere-tracer"))
    st = val_worker.stats()
    calls = int(ref.sdf_full.astype(np.int64).sum())
    assert st["sum_sdf_full"] == calls and calls % 2 == 0
    passes = calls // 2  # every scene-distance query visits both objects
    assert st["evals_by_type"] == [0, 3 * passes, passes]
    assert st["operator_flops"] > 0 and st["algorithmic_flops"] > st["operator_flops"]


def test_operator_scene_rejections():
    ctx = rb.Context(0)
    pl = sm.get_preset(11)
    t, m, q = pl.arrays()
    bad = pl.op_nodes.copy()
    bad["kind"][0] = 42
    with pytest.raises(rb.RmError) as ei:
        ctx.upload_scene(t, m, q, "BVH", op_nodes=bad, object_root=pl.object_root)
    assert ei.value.code == _lib.RM_ERR_UNSUPPORTED_PRIMITIVE
    bad = pl.op_nodes.copy()
    bad["child"][1][0] = 0
    with pytest.raises(rb.RmError) as ei:
        ctx.upload_scene(t, m, q, "BVH", op_nodes=bad, object_root=pl.object_root)
    assert ei.value.code == _lib.RM_ERR_ARG and "cycle" in str(ei.value)
    ctx.upload_scene(t, m, q, "BVH", op_nodes=pl.op_nodes, object_root=pl.object_root)  # still usable afterwards
    ctx.close()


def test_operator_full_size_properties(fast_worker, val_worker):
    """1080p, sizes the oracle is too slow for: both builds agree with each other, band partition invariance."""
    W, H = 1920, 1080
    job = make_job(W, H, 17, "BVH", "sphere-tracer", 0.15, 0.5)
    a = fast_worker.on_message(job, extras=True)
    b = val_worker.on_message(job, extras=True)
    same = (a.depth == b.depth) & (a.sdfEval == b.sdfEval)
    assert same.mean() >= PIXEL_AGREEMENT
    top = val_worker.on_message(make_job(W, H, 17, "BVH", "sphere-tracer", 0.15, 0.5, 0, 500))
    bot = val_worker.on_message(make_job(W, H, 17, "BVH", "sphere-tracer", 0.15, 0.5, 500, H))
    assert np.array_equal(np.concatenate([top.sdfEval, bot.sdfEval]), b.sdfEval)
    assert np.array_equal(np.concatenate([top.normal, bot.normal]), b.normal)


# ------------------------------------------------------------------------------------------ random operator trees
def _random_tree(rng, depth, has_twist):
    """A random operator tree over rotated / translated leaves (uses every operator kind)."""
    def leaf():
        pos = rng.uniform(-1.2, 1.2, 3)
        rot = tuple(rng.uniform(-1.0, 1.0, 3)) if rng.random() < 0.5 else None
        k = rng.integers(0, 3)
        if k == 0:
            return sm.create_sphere(*pos, rng.uniform(0.2, 0.6), rot)
        if k == 1:
            return sm.create_box(*pos, tuple(rng.uniform(0.1, 0.5, 3)), rot)
        return sm.create_torus(*pos, rng.uniform(0.3, 0.7), rot)
    if depth == 0:
        return leaf()
    k = rng.integers(0, 7)
    if k == 0:
        return sm.create_round(_random_tree(rng, depth - 1, has_twist), rng.uniform(0.01, 0.2))
    if k == 1:
        has_twist.append(True)
        return sm.create_twist(_random_tree(rng, depth - 1, has_twist), rng.uniform(0.5, 4.0))
    if k == 2:
        return sm.create_smooth_union(_random_tree(rng, depth - 1, has_twist), _random_tree(rng, depth - 1, has_twist), rng.uniform(0.0005, 0.3))
    if k == 3:
        return sm.create_smooth_subtract(_random_tree(rng, depth - 1, has_twist), _random_tree(rng, depth - 1, has_twist), rng.uniform(0.01, 0.3))
    if k == 4:
        return sm.create_repetition(_random_tree(rng, depth - 1, has_twist), tuple(rng.uniform(1.5, 3.0, 3)))
    if k == 5:
        return sm.create_animated_translate(_random_tree(rng, depth - 1, has_twist), tuple(rng.uniform(-1, 1, 3)), rng.uniform(0.2, 1.5), rng.uniform(0.001, 0.02))
    return leaf()


@pytest.mark.parametrize("seed", range(48))
def test_validation_random_operator_trees(oracle, seed):
    """Fuzz: random trees (depth <= 3, every operator kind, rotated leaves), random accel / algorithm / camera / time.
    Bit-exact against the oracle; trees containing a Twist (libm sin / cos on both sides) must at least meet the
    north-star agreement bar — in practice they are bit-exact too."""
    rng = np.random.default_rng(1000 + seed)
    has_twist = []
    objs = [_random_tree(rng, int(rng.integers(1, 4)), has_twist) for _ in range(int(rng.integers(1, 4)))]
    accel = ["None", "Octree", "BVH"][seed % 3]
    alg = ALGS[seed % len(ALGS)]
    pitch, yaw, time = float(rng.uniform(-0.8, 0.8)), float(rng.uniform(0, 6.28)), float(rng.uniform(0, 500))
    W, H = 72, 40
    pl = sm.flatten(objs)
    t, m, q = pl.arrays()
    osc = oracle.OracleScene()
    if pl.op_nodes is None:
        osc.set_prims(t, m, q)
    else:
        osc.set_tree(t, m, q, pl.op_nodes, pl.object_root)
    ref = osc.build_accel(accel).set_camera(pitch, yaw).set_time(time).render(W, H, alg)
    ctx = rb.Context(0, validate_fp64=True)
    ctx.upload_scene(t, m, q, accel, op_nodes=pl.op_nodes, object_root=pl.object_root)
    cam = Camera()
    cam.set_angles(pitch, yaw)
    f = ctx.render(rb.Context.make_request(W, H, cam.get_rotation_matrix3(), cam.get_position(), algorithm=alg, time=time), extras=True)
    ctx.close()
    if has_twist:
        same = (f.depth == ref.depth) & (f.sdfEval == ref.sdfEval) & (f.iters == ref.iters)
        assert same.mean() >= PIXEL_AGREEMENT
    else:
        assert_bit_exact(f, ref, oracle, W, H)


# ------------------------------------------------------------------------------------------ Mandelbulb
def _agree(f, ref):
    same = (f.depth == ref.depth) & (f.sdfEval == ref.sdfEval) & (f.iters == ref.iters)
    same &= (f.normal.reshape(-1, 3) == ref.normal.reshape(-1, 3)).all(1)
    return float(same.mean())


@pytest.mark.parametrize("accel", ["None", "Octree", "BVH"])
@pytest.mark.parametrize("time", [0.0, 5000.0])
def test_mandelbulb_preset_agreement(val_worker, oracle, accel, time):
    """Preset 13.  atan2 / asin / pow / log / sin / cos are libm calls on both sides (CUDA vs glibc, each <= 1-2 ulp) inside a
    chaotic iteration, so the bar is the north star's pixel agreement (>= 99.9 % of pixels identical in every plane), not
    bit-exactness; in practice the frames below differ in at most a handful of pixels."""
    W, H = 128, 72
    ref = _oracle(oracle, 13, accel, 0.3, 0.6, time=time).render(W, H, "sphere-tracer")
    f = val_worker.on_message(make_job(W, H, 13, accel, "sphere-tracer", 0.3, 0.6, time=time), extras=True)
    assert (ref.depth_f64 < 10).sum() > 500, "the fractal should be visible"
    assert np.array_equal(f.depth_f64 < 10, ref.depth_f64 < 10) or _agree(f, ref) >= PIXEL_AGREEMENT
    assert _agree(f, ref) >= PIXEL_AGREEMENT


@pytest.mark.parametrize("alg", ALGS)
def test_mandelbulb_all_algorithms(val_worker, oracle, alg):
    W, H = 96, 54
    ref = _oracle(oracle, 13, "None").render(W, H, alg)
    f = val_worker.on_message(make_job(W, H, 13, "None", alg), extras=True)
    assert _agree(f, ref) >= PIXEL_AGREEMENT


def test_mandelbulb_default_build_and_animation(fast_worker, oracle):
    W, H = 128, 72
    ref = _oracle(oracle, 13, "BVH", 0.1, 0.4).render(W, H, "sphere-tracer")
    f = fast_worker.on_message(make_job(W, H, 13, "BVH", "sphere-tracer", 0.1, 0.4), shader="phong", extras=True)
    px, dz = fast_agreement(f, ref, oracle, W, H)
    assert px >= PIXEL_AGREEMENT, f"pixel agreement {px} (depth-only among hits {dz})"
    a = fast_worker.on_message(make_job(64, 36, 13, "None", time=0.0))
    b = fast_worker.on_message(make_job(64, 36, 13, "None", time=20000.0))  # phi += time * -0.0001 per inner iteration
    assert not np.array_equal(a.depth, b.depth)


def test_time_does_not_reach_below_an_animated_translate(oracle):
    """AnimatedTranslate.setTime keeps the time for itself (animatedTranslate.ts:30-32): an AnimatedTranslate or a Mandelbulb
    underneath stays at time 0 whatever Job.time says.  (Found by the random-tree test.)"""
    W, H = 80, 45
    inner = sm.create_animated_translate(sm.create_sphere(0.2, 0.1, 0.0, 0.5), (0, 1, 0), 1.5, 0.01)
    objs = [sm.create_animated_translate(inner, (1, 0, 0), 1.0, 0.004),
            sm.create_animated_translate(sm.create_mandelbulb(-1.5, 0, 0, 8, 10, True, -0.01), (0, 0, 1), 0.5, 0.003)]
    pl = sm.flatten(objs)
    t, m, q = pl.arrays()
    for time in (0.0, 321.0):
        ref = oracle.OracleScene().set_tree(t, m, q, pl.op_nodes, pl.object_root).build_accel("None").set_camera(0.1, 0.2).set_time(time).render(W, H)
        ctx = rb.Context(0, validate_fp64=True)
        ctx.upload_scene(t, m, q, "None", op_nodes=pl.op_nodes, object_root=pl.object_root)
        cam = Camera()
        cam.set_angles(0.1, 0.2)
        f = ctx.render(rb.Context.make_request(W, H, cam.get_rotation_matrix3(), cam.get_position(), time=time), extras=True)
        ctx.close()
        assert _agree(f, ref) >= PIXEL_AGREEMENT, time
        assert (ref.depth_f64 < 10).sum() > 100


def test_mandelbulb_inside_an_operator_tree(oracle):
    """A Mandelbulb leaf under Round + SmoothUnion: the leaf evaluator is shared with the tree interpreter."""
    W, H = 96, 54
    objs = [sm.create_smooth_union(sm.create_round(sm.create_mandelbulb(0.4, 0, 0, 8, 12, True, -0.0001), 0.02),
                                   sm.create_sphere(-1.0, 0.2, 0.0, 0.4), 0.1)]
    pl = sm.flatten(objs)
    t, m, q = pl.arrays()
    ref = oracle.OracleScene().set_tree(t, m, q, pl.op_nodes, pl.object_root).build_accel("Octree").set_camera(0.0, 0.3).set_time(123.0).render(W, H)
    ctx = rb.Context(0, validate_fp64=True)
    ctx.upload_scene(t, m, q, "Octree", op_nodes=pl.op_nodes, object_root=pl.object_root)
    cam = Camera()
    cam.set_angles(0.0, 0.3)
    f = ctx.render(rb.Context.make_request(W, H, cam.get_rotation_matrix3(), cam.get_position(), time=123.0), extras=True)
    assert _agree(f, ref) >= PIXEL_AGREEMENT
    ctx.close()
