"""The Node N-API addon (addon/rm_napi.cc) EXECUTED end to end on the GPU under the mock Node-API runtime of tests/napi_mock.cc
(run with -m gpu).  No Node.js exists in this image; the mock implements the N-API calls the addon uses, runs the async work on
real threads and enforces handle scopes, so the flow ts/gpuWorkerShim.ts drives from JavaScript — uploadScene once, the row-band
Jobs of a frame posted together (main.ts:444-486), one promise per Job, Result typed arrays over page-locked external
ArrayBuffers, finalizers returning the blocks — runs for real and its frames are compared with the ctypes path bit for bit."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import make_job

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = tmp_path_factory.mktemp("addon") / "addon_harness"
    cxx = os.environ.get("CXX", "g++")
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-Wall", "-pthread", os.path.join(ROOT, "tests", "addon_harness.cc"),
                           os.path.join(ROOT, "tests", "napi_mock.cc"), os.path.join(ROOT, "addon", "rm_napi.cc"), "-o", str(out),
                           "-L" + os.path.join(ROOT, "cpu_raymarcher_b200"), "-lrm_b200", "-Wl,-rpath," + os.path.join(ROOT, "cpu_raymarcher_b200")])
    return str(out)


def _scene_file(path, preset, accel, synthetic, pitch, yaw):
    from cpu_raymarcher_b200.renderer import OP_NODE_DTYPE
    from cpu_raymarcher_b200.scene import Scene
    sc = Scene(accel)
    if synthetic:
        sc.load_synthetic(*synthetic)
    else:
        sc.load_preset(preset)
    t, m, q = sc.primitives.arrays()
    ops, roots = sc.primitives.op_nodes, sc.primitives.object_root
    has_tree = roots is not None and len(roots) > 0
    ops_b = np.ascontiguousarray(ops, OP_NODE_DTYPE).tobytes() if has_tree else b""
    roots_b = np.ascontiguousarray(roots, np.int32).tobytes() if has_tree else b""
    sc.camera.set_angles(pitch, yaw)
    with open(path, "wb") as fh:
        fh.write(struct.pack("4i", len(t), len(ops_b), len(roots) if has_tree else 0, 0))
        fh.write(np.ascontiguousarray(t, np.uint8).tobytes() + np.ascontiguousarray(m, np.float32).tobytes() + np.ascontiguousarray(q, np.float64).tobytes())
        fh.write(ops_b + roots_b)
        fh.write(np.asarray(sc.camera.get_rotation_matrix3(), np.float32).tobytes() + np.asarray(sc.camera.get_position(), np.float32).tobytes())


CASES = [
    dict(W=320, H=182, preset=2, accel="BVH", alg="sphere-tracer", bands=4, env={}),
    dict(W=200, H=120, preset=16, accel="BVH", alg="adaptive-step-v3", bands=3, env={}),                     # operator tree (Screw)
    dict(W=256, H=150, preset=1, accel="BVH", alg="sphere-tracer", bands=4, synthetic=(3000, 0x5EED0001), env={"RM_DEVICES": "0,0"}),
    dict(W=1920, H=1080, preset=3, accel="Octree", alg="sphere-tracer", bands=4, env={}),                    # big enough for the early band download
    dict(W=160, H=90, preset=9, accel="None", alg="fixed-step", bands=4, env={"NAPI_MOCK_NO_EXTERNAL_BUFFERS": "1"}),  # plain-ArrayBuffer fallback
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"p{c['preset']}-{c['accel']}-{c['bands']}bands" + ("-" + "+".join(c["env"]) if c["env"] else ""))
def test_addon_frames_equal_ctypes_frames(harness, fast_worker, tmp_path, case):
    W, H = case["W"], case["H"]
    scene, out = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    _scene_file(scene, case["preset"], case["accel"], case.get("synthetic"), 0.15, 0.6)
    env = dict(os.environ, **case["env"])
    r = subprocess.run([harness, scene, out, str(W), str(H), case["alg"], case["accel"], str(case["bands"])], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-2000:]
    assert "OK" in r.stdout
    want_dev = 2 if case["env"].get("RM_DEVICES") else None
    if want_dev:
        assert f"devices {want_dev}" in r.stdout
    ref = fast_worker.on_message(make_job(W, H, case["preset"], case["accel"], case["alg"], 0.15, 0.6, synthetic=case.get("synthetic")))
    rs = fast_worker.stats()
    raw = np.fromfile(out, np.uint8)
    n = W * H
    frame_bytes = 8 * n
    assert raw.size == 2 * frame_bytes + 48
    for k in range(2):  # the second frame was written into recycled page-locked blocks
        fr = raw[k * frame_bytes:(k + 1) * frame_bytes]
        assert np.array_equal(fr[:n], ref.depth), k
        assert np.array_equal(fr[n:4 * n], ref.normal), k
        assert np.array_equal(fr[4 * n:6 * n].view(np.uint16), ref.sdfEval), k
        assert np.array_equal(fr[6 * n:8 * n].view(np.uint16), ref.iters), k
    st = raw[2 * frame_bytes:].view(np.float64)
    assert (st[0], st[1], st[2], st[3], st[4]) == (rs["sum_sdf"], rs["max_sdf"], rs["min_sdf"], rs["sum_iters"], n)
    fin = int(r.stdout.split("finalized_external_buffers")[1].split()[0])
    if "NAPI_MOCK_NO_EXTERNAL_BUFFERS" in case["env"]:
        assert fin == 0
    else:
        n_jobs = sum(1 for i in range(case["bands"]) if min(i * -(-H // case["bands"]), H) < min((i + 1) * -(-H // case["bands"]), H))
        assert fin == 2 * n_jobs, "every Result's external ArrayBuffer must hand its block back"
