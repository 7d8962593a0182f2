"""CPU tests of the host side: C-ABI exports, loud failure without a GPU, native builders and the Python
mirrors of sceneManager.ts / camera.ts against the oracle (independent implementations, bit-exact)."""
import ctypes as C
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

import cpu_raymarcher_b200 as rb
from cpu_raymarcher_b200 import _lib, multigpu
from cpu_raymarcher_b200 import scene_manager as sm
from cpu_raymarcher_b200.camera import Camera

from conftest import HAVE_GPU

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "rm.h")).read()
    declared = set(re.findall(r"^\s*(?:int|void|const char\*)\s+(rm_[a-z0-9_]+)\s*\(", hdr, re.M))
    assert declared, "no declarations parsed"
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/rm.h but not exported"
    assert declared == set(_lib.EXPORTS)
    assert L.rm_abi_version() == 3


def test_struct_layouts_match_c_abi():
    assert C.sizeof(_lib.BvhNode) == 40
    assert C.sizeof(_lib.OctreeNode) == 48
    assert _lib.OctreeNode.min_distance.offset == 40
    assert C.sizeof(_lib.Request) % 8 == 0
    assert C.sizeof(_lib.OpNode) == 128 == sm.OP_NODE_DTYPE.itemsize
    assert _lib.OpNode.transform.offset == 64 and _lib.OpNode.p.offset == 16
    assert _lib.Scene.object_root.offset == C.sizeof(_lib.Scene) - 8


@pytest.mark.skipif(HAVE_GPU, reason="only meaningful on a machine without a GPU")
def test_no_gpu_fails_loudly_no_cpu_fallback():
    with pytest.raises(rb.RmError) as ei:
        rb.Context()
    assert ei.value.code == _lib.RM_ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cpu_raymarcher_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "oracle/" not in txt.replace("not share sources with oracle/", ""), f


@pytest.mark.parametrize("idx", sm.SUPPORTED_PRESETS)
def test_presets_match_oracle_bitwise(oracle, idx):
    pl = sm.get_preset(idx)
    t, m, q = pl.arrays()
    if idx in sm.OPERATOR_PRESETS:  # operator trees: leaves + the flat node array, same pre-order on both sides
        ot, om, oq, onodes, oroots = oracle.OracleScene().load_preset(idx).get_tree()
        assert pl.op_nodes.tobytes() == onodes.tobytes()
        assert np.array_equal(pl.object_root, oroots)
    else:
        ot, om, oq = oracle.OracleScene().load_preset(idx).get_prims()
        assert pl.op_nodes is None
    assert np.array_equal(t, ot)
    assert np.array_equal(bits(m), bits(om))
    assert np.array_equal(bits(q), bits(oq))


def test_every_reference_preset_is_supported():
    assert sm.SUPPORTED_PRESETS == tuple(range(19)) == tuple(range(sm.get_preset_count()))
    t, m, q = sm.get_preset(13).arrays()  # Mandelbulb [A]: a leaf primitive with a scaled transform (sceneManager.ts:62-63)
    assert list(t) == [sm.MANDELBULB] and list(q[0]) == [8.0, 80.0, 1.0, -0.0001] and m[0][0] == 0.5
    with pytest.raises(IndexError):
        sm.get_preset(19)


def test_synthetic_scene_matches_oracle_bitwise(oracle):
    t, m, q = sm.synthetic_spheres(5000).arrays()
    ot, om, oq = oracle.OracleScene().load_synthetic(5000).get_prims()
    assert np.array_equal(t, ot) and np.array_equal(bits(m), bits(om)) and np.array_equal(bits(q), bits(oq))
    # first 7 primitives are the reference's "Random Spheres"
    t1, m1, q1 = sm.get_preset(1).arrays()
    assert np.array_equal(bits(m[:7]), bits(m1)) and np.array_equal(q[:7], q1)
    # vectorised transform == scalar transform (incl. the sign of zeros)
    c = np.array([[0.25, -1.5, 2.0], [-2.4, 0.0, 0.1]])
    batch = sm.get_transform_batch(c)
    for k in range(2):
        assert np.array_equal(bits(np.array(sm.get_transform(*c[k]), np.float32)), bits(batch[k]))


@pytest.mark.parametrize("pitch,yaw", [(0, 0), (0.3, 1.1), (-2.0, 5.4), (0.0, 0.015 * 77), (1.2, -3.0)])
def test_camera_matches_oracle_bitwise(oracle, pitch, yaw):
    c = Camera()
    c.set_angles(pitch, yaw)
    r, o = oracle.OracleScene().set_camera(pitch, yaw).get_camera()
    assert np.array_equal(bits(r), bits(c.get_rotation_matrix3()))
    assert np.array_equal(bits(o), bits(c.get_position()))


def test_camera_analytics_rotation_accumulates_like_main_ts(oracle):
    c = Camera()
    s = oracle.OracleScene()
    for _ in range(25):  # main.ts:438-441: yaw += 0.015 before every frame
        c.rotate_camera(0, 0.015)
        s.rotate_camera(0, 0.015)
    assert c.get_angles() == s.get_angles()
    r, o = s.get_camera()
    assert np.array_equal(bits(r), bits(c.get_rotation_matrix3()))


BVH_DT = np.dtype([("bmin", "<f4", 3), ("bmax", "<f4", 3), ("l", "<i4"), ("r", "<i4"), ("pf", "<i4"), ("pc", "<i4")])
OCT_DT = np.dtype([("bmin", "<f4", 3), ("bmax", "<f4", 3), ("fc", "<i4"), ("pf", "<i4"), ("pc", "<i4"), ("lvl", "u1"), ("emp", "u1"),
                   ("pad", "u1", 2), ("md", "<f8")])


def _scene_cases(big=()):
    # 20 000 and 100 000 objects: past the sizes where the builders fork threads (BVH subtrees, leaf-grid slabs, octree
    # minDistance queries) and where the octree's nearest-box search prunes instead of trying every box
    return [("preset", i) for i in sm.SUPPORTED_PRESETS] + [("synthetic", n) for n in (300, 3000, 20000) + tuple(big)]


@pytest.mark.parametrize("kind,arg", _scene_cases(big=(100000,)))
def test_native_bvh_builder_matches_oracle(oracle, kind, arg):
    if kind == "preset":
        pl = sm.get_preset(arg)
        osc = oracle.OracleScene().load_preset(arg)
    else:
        pl = sm.synthetic_spheres(arg)
        osc = oracle.OracleScene().load_synthetic(arg)
    t, m, q = pl.arrays()
    ob, ol, oleaf = osc.build_accel("BVH").bvh_flat()
    nodes, nn, leaf = rb.build_bvh_scene(t, m, q, pl.op_nodes, pl.object_root)
    if pl.op_nodes is None:  # the array form of the entry point gives the same bytes
        n2, nn2, leaf2 = rb.build_bvh(t, m, q)
        assert nn2 == nn and bytes(n2) == bytes(nodes) and np.array_equal(leaf, leaf2)
    a = np.frombuffer(nodes, dtype=BVH_DT, count=nn)
    assert nn == len(ob)
    assert np.array_equal(bits(a["bmin"]), bits(ob[:, :3])) and np.array_equal(bits(a["bmax"]), bits(ob[:, 3:]))
    assert np.array_equal(a["l"], ol[:, 0]) and np.array_equal(a["r"], ol[:, 1])
    assert np.array_equal(a["pf"], ol[:, 2]) and np.array_equal(a["pc"], ol[:, 3])
    assert np.array_equal(leaf, oleaf)
    # every scene object lives in exactly one leaf (so the reference's Set de-duplication never matters)
    assert sorted(leaf.tolist()) == list(range(pl.n_objects))


@pytest.mark.parametrize("kind,arg", _scene_cases())
def test_native_octree_builder_matches_oracle(oracle, kind, arg):
    if kind == "preset":
        pl = sm.get_preset(arg)
        osc = oracle.OracleScene().load_preset(arg)
    else:
        pl = sm.synthetic_spheres(arg)
        osc = oracle.OracleScene().load_synthetic(arg)
    t, m, q = pl.arrays()
    ob, ol, lvl, emp, mind, oleaf = osc.build_accel("Octree").octree_flat()
    nodes, nn, leaf = rb.build_octree_scene(t, m, q, pl.op_nodes, pl.object_root)
    if pl.op_nodes is None:
        n2, nn2, leaf2 = rb.build_octree(t, m, q)
        assert nn2 == nn and bytes(n2) == bytes(nodes) and np.array_equal(leaf, leaf2)
    a = np.frombuffer(nodes, dtype=OCT_DT, count=nn)
    assert nn == len(ob)
    assert np.array_equal(bits(a["bmin"]), bits(ob[:, :3])) and np.array_equal(bits(a["bmax"]), bits(ob[:, 3:]))
    assert np.array_equal(a["fc"], ol[:, 0]) and np.array_equal(a["pf"], ol[:, 1]) and np.array_equal(a["pc"], ol[:, 2])
    assert np.array_equal(a["lvl"], lvl) and np.array_equal(a["emp"], emp)
    assert np.array_equal(bits(a["md"]), bits(mind))
    assert np.array_equal(leaf, oleaf)


def test_builder_entry_points_reject_bad_arguments():
    """rm_build_bvh / rm_build_octree: the sizing call (nodes == NULL) reports the counts, a too-small capacity and missing
    arrays are RM_ERR_ARG — never a partial write."""
    L = _lib.lib()
    t, m, q = sm.synthetic_spheres(50).arrays()
    t, m, q = np.ascontiguousarray(t, np.uint8), np.ascontiguousarray(m, np.float32), np.ascontiguousarray(q, np.float64)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    for fn, node_t in ((L.rm_build_bvh, _lib.BvhNode), (L.rm_build_octree, _lib.OctreeNode)):
        nn, nl = C.c_int32(0), C.c_int32(0)
        assert fn(50, ptr(t), ptr(m), ptr(q), 0, None, C.byref(nn), None, C.byref(nl)) == 0
        assert nn.value > 0 and nl.value >= 50
        need_n, need_l = nn.value, nl.value
        nodes = (node_t * need_n)()
        leaf = np.full(need_l, -7, np.int32)
        small = C.c_int32(need_n - 1)
        assert fn(50, ptr(t), ptr(m), ptr(q), 0, nodes, C.byref(small), ptr(leaf), C.byref(nl)) == _lib.RM_ERR_ARG
        assert (leaf == -7).all(), "nothing may be written when the capacity is too small"
        small_l = C.c_int32(need_l - 1)
        nn.value = need_n
        assert fn(50, ptr(t), ptr(m), ptr(q), 0, nodes, C.byref(nn), ptr(leaf), C.byref(small_l)) == _lib.RM_ERR_ARG
        assert fn(50, ptr(t), ptr(m), ptr(q), 0, nodes, C.byref(nn), None, C.byref(nl)) == _lib.RM_ERR_ARG  # leaf array missing
        assert fn(50, None, ptr(m), ptr(q), 0, None, C.byref(nn), None, C.byref(nl)) == _lib.RM_ERR_ARG
        assert fn(-1, ptr(t), ptr(m), ptr(q), 0, None, C.byref(nn), None, C.byref(nl)) == _lib.RM_ERR_ARG
        assert fn(50, ptr(t), ptr(m), ptr(q), 0, None, None, None, C.byref(nl)) == _lib.RM_ERR_ARG
        nn.value, nl.value = need_n, need_l
        assert fn(50, ptr(t), ptr(m), ptr(q), 0, nodes, C.byref(nn), ptr(leaf), C.byref(nl)) == 0
        assert (nn.value, nl.value) == (need_n, need_l) and (leaf >= 0).all() and (leaf < 50).all()


def test_leaf_grid_lists_equal_brute_force(tmp_path):
    """The leaf grid of the fast BVH path (csrc/rm_build.cpp build_leaf_grid: per-cell leaf lists + the six direction lists of the
    DDA walk, filled by z-slab threads) against a brute-force rebuild of every list, in a C++ harness linked with the product's
    builder source: above and below the size where the threads start, a flat scene (one z layer), one sphere, no sphere."""
    cxx = shutil.which(os.environ.get("CXX", "g++"))
    if not cxx:
        pytest.skip("no C++ compiler")
    exe = tmp_path / "leafgrid_harness"
    inc = [p for p in ("/usr/local/cuda/include",) if os.path.isdir(p)]
    subprocess.check_call([cxx, "-std=c++17", "-O2", "-Wall", "-ffp-contract=off", "-pthread"] + ["-I" + p for p in inc] +
                          [os.path.join(ROOT, "tests", "leafgrid_harness.cc"), os.path.join(ROOT, "cpu_raymarcher_b200", "csrc", "rm_build.cpp"), "-o", str(exe)])
    for args in (["9000", "0.01", "0.05", "5"], ["9000", "0.01", "0.05", "6", "flat"], ["500", "0.1", "0.6", "7"], ["1", "0.1", "0.6", "8"], ["0", "0.1", "0.6", "9"]):
        r = subprocess.run([str(exe)] + args, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and r.stdout.startswith("OK"), (args, r.stdout, r.stderr)


def test_builders_handle_empty_and_rotated_scenes(oracle):
    nodes, nn, leaf = rb.build_bvh(np.zeros(0, np.uint8), np.zeros((0, 16), np.float32), np.zeros((0, 4)))
    assert nn == 1 and len(leaf) == 0
    pl = sm.PrimitiveList()
    sm.add_box(pl, 0.3, -0.2, 0.5, (0.4, 0.2, 0.7), rotation=(0.3, 1.1, -0.8))
    sm.add_torus(pl, -1.0, 0.4, 0.0, 0.8, rotation=(1.0, 0.0, 0.5))
    sm.add_sphere(pl, 1.0, 1.0, -1.0, 0.3)
    sm.add_box(pl, -0.6, -0.9, 0.2, (0.1, 0.5, 0.3))
    sm.add_sphere(pl, 0.0, 0.2, 1.4, 0.25, rotation=(0.1, 0.2, 0.3))
    t, m, q = pl.arrays()
    osc = oracle.OracleScene().set_prims(t, m, q).build_accel("BVH")
    ob, ol, oleaf = osc.bvh_flat()
    nodes, nn, leaf = rb.build_bvh(t, m, q)
    a = np.frombuffer(nodes, dtype=BVH_DT, count=nn)
    assert np.array_equal(bits(a["bmin"]), bits(ob[:, :3])) and np.array_equal(leaf, oleaf)
    # transforms with rotation agree with the oracle's getTransform
    out = np.zeros(16, np.float32)
    rot = np.array([0.3, 1.1, -0.8], np.float32)
    oracle.lib().orc_get_transform(0.3, -0.2, 0.5, rot.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert np.array_equal(bits(out), bits(m[0]))


def test_stripe_partition_covers_every_row_once():
    for H in (1, 7, 8, 9, 63, 64, 2160):
        for world in (1, 2, 3, 4, 8):
            rows = np.concatenate([multigpu.stripe_rows_of(r, world, H) for r in range(world)])
            assert sorted(rows.tolist()) == list(range(H))


def test_plane_layout_is_aligned_and_disjoint():
    lay = multigpu.plane_layout(1920, 1080)
    n = 1920 * 1080
    sizes = dict(depth=n, normal=3 * n, sdf=2 * n, iters=2 * n, rgba=4 * n)
    spans = sorted((lay[k], lay[k] + sizes[k]) for k in sizes)
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0
    assert all(lay[k] % 256 == 0 for k in sizes) and lay["total"] >= spans[-1][1]


def test_napi_addon_compiles_against_the_c_abi(tmp_path):
    """addon/rm_napi.cc (the Node binding of INTEGRATION.md) must stay well-formed against include/rm.h; Node.js itself is
    not available here, so it is compiled against the restated N-API prototypes of addon/napi_min.h."""
    import shutil
    import subprocess
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    out = tmp_path / "rm_napi.o"
    subprocess.check_call([cxx, "-std=c++17", "-fPIC", "-Wall", "-c", os.path.join(ROOT, "addon", "rm_napi.cc"), "-o", str(out)])
    assert out.stat().st_size > 0


def test_header_is_plain_c(tmp_path):
    """include/rm.h is the drop-in boundary: it must compile as C (no C++-isms, no torch / CUDA types)."""
    import shutil
    import subprocess
    cc = shutil.which("gcc")
    if cc is None:
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text('#include "rm.h"\nint main(void) { rm_scene s; rm_op_node n; (void)s; (void)n; return sizeof(rm_stats_t) > 0 ? 0 : 1; }\n')
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")])


@pytest.mark.skipif(HAVE_GPU, reason="only meaningful on a machine without a GPU")
def test_pool_without_gpu_fails_loudly_and_addon_harness_reports_it(tmp_path):
    """rm_pool_create has no CPU fallback either; and the N-API addon (run here under the mock runtime of tests/napi_mock.cc)
    turns that failure into a thrown JS error instead of a hang or a silent empty frame."""
    with pytest.raises(rb.RmError) as ei:
        rb.RaymarchPool()
    assert ei.value.code == _lib.RM_ERR_CUDA and "no CPU fallback" in str(ei.value)
    cxx = shutil.which(os.environ.get("CXX", "g++"))
    if not cxx:
        pytest.skip("no C++ compiler")
    exe = tmp_path / "addon_harness"
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-Wall", "-pthread", os.path.join(ROOT, "tests", "addon_harness.cc"),
                           os.path.join(ROOT, "tests", "napi_mock.cc"), os.path.join(ROOT, "addon", "rm_napi.cc"), "-o", str(exe),
                           "-L" + os.path.join(ROOT, "cpu_raymarcher_b200"), "-lrm_b200", "-Wl,-rpath," + os.path.join(ROOT, "cpu_raymarcher_b200")])
    import struct
    scene = tmp_path / "scene.bin"
    scene.write_bytes(struct.pack("4i", 1, 0, 0, 0) + bytes([0]) + np.eye(4, dtype=np.float32).tobytes() + np.array([1.5, 0, 0, 0]).tobytes() +
                      np.eye(3, dtype=np.float32).tobytes() + np.array([0, 0, 3], np.float32).tobytes())
    r = subprocess.run([str(exe), str(scene), str(tmp_path / "out.bin"), "32", "32", "sphere-tracer", "None", "4"], capture_output=True, text=True)
    assert r.returncode == 1 and "uploadScene threw: RM_ERR_CUDA" in r.stderr and "no CPU fallback" in r.stderr
