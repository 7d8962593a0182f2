"""CPU tests of the oracle itself: hand-derived known-answer vectors, the JS numeric
model (Appendix A), reference-guaranteed invariants and the committed golden fixtures.
The reference ships no tests or golden vectors (package.json:7) -> parity is unpinned by the reference;
these are the pins this repo creates."""
import math
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_kat1_centre_pixel(oracle):
    # config 1, 512x512, yaw 0, pixel (256,256): t=1.5 after 2 steps; normal bytes (127,127,255); depth byte 2
    f = oracle.OracleScene().load_preset(0).render(512, 512, nthreads=4, y_start=256, y_end=257)
    i = 256
    assert f.depth_f64[i] == 1.5
    assert f.depth[i] == 2  # 1.5 rounds half-to-even
    assert list(f.normal[3 * i:3 * i + 3]) == [127, 127, 255]
    assert f.sdfEval[i] == 6 and f.iters[i] == 2
    heat = oracle.shade("sdf-heatmap", f.depth, f.normal, f.sdfEval, f.iters, 512, 1)
    assert list(heat[4 * i:4 * i + 4]) == [60, 255, 0, 255]
    heat = oracle.shade("iteration-heatmap", f.depth, f.normal, f.sdfEval, f.iters, 512, 1)
    assert list(heat[4 * i:4 * i + 4]) == [20, 255, 0, 255]


def test_kat2_miss_pixel(oracle):
    f = oracle.OracleScene().load_preset(0).render(512, 512, nthreads=4, y_start=0, y_end=1)
    assert f.depth_f64[0] > 10 and f.depth_f64[0] <= 20
    assert list(f.normal[0:3]) == [128, 128, 128]
    assert f.sdfEval[0] == f.iters[0]  # N = 1, no normal taps on a miss
    assert f.depth[0] == int(np.round(f.depth_f64[0]))  # no tie here


def test_kat2_fixed_and_adaptive_return_exactly_10_on_miss(oracle):
    for alg in ("fixed-step", "adaptive-step"):
        f = oracle.OracleScene().load_preset(0).render(64, 64, alg, y_start=0, y_end=1)
        assert f.depth_f64[0] == 10.0 and f.depth[0] == 10


def test_kat3_bvh_miss_has_zero_counters(oracle):
    f = oracle.OracleScene().load_preset(2).build_accel("BVH").render(128, 72)
    miss = (f.sdf_full == 0)
    assert miss.any()
    assert np.all(f.depth_f64[miss] == 10.0) and np.all(f.iters[miss] == 0)
    st = oracle.stats(f.sdfEval, f.iters)
    assert st["min_sdf"] == 0


@pytest.mark.parametrize("preset,n", [(0, 1), (1, 7), (4, 7), (8, 2), (9, 3)])
def test_invariant_no_accel_sphere_tracer(oracle, preset, n):
    f = oracle.OracleScene().load_preset(preset).render(96, 64)
    hit = f.depth_f64 < 10
    assert hit.any() and (~hit).any()
    assert np.all(f.sdf_full[hit] == n * (f.iters_full[hit] + 4))
    assert np.all(f.sdf_full[~hit] == n * f.iters_full[~hit])


def test_tile_partition_invariance(oracle):
    s = oracle.OracleScene().load_preset(3).build_accel("Octree")
    full = s.render(80, 60, "adaptive-step-v3")
    parts = [s.render(80, 60, "adaptive-step-v3", y_start=a, y_end=b) for a, b in ((0, 15), (15, 30), (30, 45), (45, 60))]
    assert np.array_equal(np.concatenate([p.sdfEval for p in parts]), full.sdfEval)
    assert np.array_equal(np.concatenate([p.normal for p in parts]), full.normal)
    assert np.array_equal(np.concatenate([p.depth for p in parts]), full.depth)


def test_hypot_is_v8_kahan(oracle):
    L = oracle.lib()
    assert L.orc_hypot3(3.0, 4.0, 0.0) == 5.0
    assert L.orc_hypot3(0.0, 0.0, 0.0) == 0.0
    assert math.isinf(L.orc_hypot3(float("inf"), float("nan"), 1.0))  # Infinity wins over NaN
    assert math.isnan(L.orc_hypot3(1.0, float("nan"), 1.0))
    # max-scaled: no overflow where the naive sum of squares would
    assert L.orc_hypot3(1e200, 1e200, 0.0) == pytest.approx(math.sqrt(2) * 1e200, rel=1e-15)


def test_to_uint8_clamp(oracle):
    L = oracle.lib()
    cases = {0.5: 0, 1.5: 2, 2.5: 2, 127.5: 128, 254.5: 254, 255.5: 255, -3.0: 0, 300.0: 255, 1.4999: 1, 1.5001: 2}
    for x, want in cases.items():
        assert L.orc_to_u8(x) == want, x
    assert L.orc_to_u8(float("nan")) == 0


def test_js_min_max(oracle):
    L = oracle.lib()
    assert math.isnan(L.orc_min2(1.0, float("nan"))) and math.isnan(L.orc_max2(float("nan"), 1.0))
    assert math.copysign(1, L.orc_min2(0.0, -0.0)) == -1 and math.copysign(1, L.orc_max2(-0.0, 0.0)) == 1
    assert L.orc_min2(2.0, 3.0) == 2.0 and L.orc_max2(2.0, 3.0) == 3.0


def test_mulberry32_known_values(oracle):
    # independent pure-python evaluation of the generator (32-bit integer ops only)
    from cpu_raymarcher_b200.scene_manager import mulberry32
    rng = mulberry32(0x5EED0001)
    py = [rng() for _ in range(8)]
    for k, v in enumerate(py):
        assert oracle.lib().orc_mulberry32(0x5EED0001, k) == v
        assert 0.0 <= v < 1.0


def test_u16_counter_wraps(oracle):
    # 70 000 spheres, no accel: every query costs 70 000 evaluations -> the Uint16Array counter wraps
    s = oracle.OracleScene().load_synthetic(70000)
    f = s.render(4, 2)
    assert np.all(f.sdf_full >= 70000)
    assert np.array_equal(f.sdfEval, (f.sdf_full % 65536).astype(np.uint16))


def test_phong_background_branch_never_fires(oracle):
    # depth is stored in world units (<= ~20), so `depth >= 255` (phongModel.ts:34) is dead: misses are lit
    f = oracle.OracleScene().load_preset(0).render(64, 64)
    rgba = oracle.shade("phong", f.depth, f.normal, f.sdfEval, f.iters, 64, 64).reshape(-1, 4)
    assert not np.any(np.all(rgba[:, :3] == [10, 10, 20], axis=1))
    assert np.all(rgba[:, 3] == 255)


def test_length_mode_switch_changes_only_ulps(oracle):
    s = oracle.OracleScene().load_preset(1)
    a = s.render(64, 48)
    oracle.lib().orc_set_length_mode(0)
    try:
        b = s.render(64, 48)
    finally:
        oracle.lib().orc_set_length_mode(1)
    assert np.array_equal(a.depth_f64 < 10, b.depth_f64 < 10)
    assert np.max(np.abs(a.depth_f64 - b.depth_f64)) < 1e-9


def _golden_cases():
    import json
    with open(os.path.join(GOLDEN, "manifest.json")) as fh:
        return json.load(fh)["cases"]


@pytest.mark.parametrize("case", _golden_cases(), ids=lambda c: c["name"])
def test_oracle_matches_committed_golden(oracle, case):
    """The golden fixtures were produced by tools/make_golden.py from this oracle; this pins the oracle against
    accidental change (and is the same data the GPU validation build is checked against)."""
    g = np.load(os.path.join(GOLDEN, case["name"] + ".npz"))
    s = oracle.OracleScene()
    if case.get("synthetic"):
        s.load_synthetic(*case["synthetic"])
    else:
        s.load_preset(case["preset"])
    s.build_accel(case["accel"]).set_camera(case["pitch"], case["yaw"]).set_time(case.get("time", 0.0))
    f = s.render(case["W"], case["H"], case["alg"], step_size=case["step"], overshoot=case["over"])
    for k in ("depth", "normal", "sdfEval", "iters"):
        assert np.array_equal(getattr(f, k), g[k]), k
    assert np.array_equal(f.depth_f64.view(np.uint64), g["depth_f64"].view(np.uint64))
