"""CPU tests of the SDF-operator widening: oracle known answers for the six operators of
src/util/primitive_operations/, the native builders over operator trees against the oracle (bit-exact), and the
structural validation of rm_op_node arrays.  No GPU needed."""
import ctypes as C
import math

import numpy as np
import pytest

import cpu_raymarcher_b200 as rb
from cpu_raymarcher_b200 import _lib
from cpu_raymarcher_b200 import scene_manager as sm


def f32(x):
    return float(np.float32(x))


def _oracle_from(oracle, objects):
    pl = sm.flatten(objects)
    t, m, q = pl.arrays()
    if pl.op_nodes is None:  # a plain primitive list
        return oracle.OracleScene().set_prims(t, m, q), pl
    return oracle.OracleScene().set_tree(t, m, q, pl.op_nodes, pl.object_root), pl


# ------------------------------------------------------------------------------------------ known answers
def test_js_round_matches_ecmascript(oracle):
    L = oracle.lib()
    for x, want in [(0.5, 1.0), (-0.5, -0.0), (2.5, 3.0), (-2.5, -2.0), (0.49999999999999994, 0.0), (-0.2, -0.0),
                    (1.4, 1.0), (-1.6, -2.0), (4503599627370497.0, 4503599627370497.0)]:
        got = L.orc_js_round(x)
        assert got == want and math.copysign(1, got) == math.copysign(1, want), x


def test_round_kat(oracle):
    """round.ts:23: box of half 0.4 rounded by 0.3, on the +x axis: (2 - f32(0.4)) stored as f32, minus 0.3."""
    s, _ = _oracle_from(oracle, [sm.create_round(sm.create_box(0, 0, 0, (0.4, 0.4, 0.4)), 0.3)])
    d, n = s.distance([2.0, 0.0, 0.0])
    assert n == 1
    assert d == f32(2.0 - f32(0.4)) - 0.3


def test_smooth_union_kat(oracle):
    """smoothUnion.ts:31-34 with d1 == d2 == d: h = k, result = d - k*k*0.25/k (k = 4 * smoothness)."""
    a = sm.create_sphere(0, 0, 0, 1.0)
    s, _ = _oracle_from(oracle, [sm.create_smooth_union(a, sm.create_sphere(0, 0, 0, 1.0), 0.2)])
    d, _ = s.distance([3.0, 0.0, 0.0])
    k = 0.2 * 4.0
    assert d == 2.0 - k * k * 0.25 / k
    # far apart: h clamps to 0 and the union is the hard min
    s2, _ = _oracle_from(oracle, [sm.create_smooth_union(sm.create_sphere(-4, 0, 0, 0.5), sm.create_sphere(4, 0, 0, 0.5), 0.1)])
    assert s2.distance([-2.0, 0.0, 0.0])[0] == 1.5


def test_smooth_subtraction_kat(oracle):
    """smoothSubstraction.ts:30-33: max(d1, -d2) + h*h*0.25/k with h = max(k - |d1 + d2|, 0)."""
    s, _ = _oracle_from(oracle, [sm.create_smooth_subtract(sm.create_sphere(0, 0, 0, 2.0), sm.create_sphere(0, 0, 0, 1.0), 0.05)])
    d1, d2 = 1.5 - 2.0, 1.5 - 1.0
    k = 0.05 * 4.0
    h = max(k - abs(d1 + d2), 0.0)
    assert s.distance([1.5, 0.0, 0.0])[0] == max(d1, -d2) + h * h * 0.25 / k


def test_repetition_kat(oracle):
    """repetition.ts:21-26: q = p - s*round(p/s); a lattice point of the 1.5 grid is a sphere centre."""
    s, _ = _oracle_from(oracle, [sm.create_repetition(sm.create_sphere(0, 0, 0, 0.3), (1.5, 1.5, 1.5))])
    assert s.distance([1.5, -3.0, 4.5])[0] == -0.3
    assert s.distance([0.75, 0.0, 0.0])[0] == 0.75 - 0.3  # the tie rounds up: q.x = 0.75 - 1.5 = -0.75
    assert s.object_geometry(0)[1] == math.inf          # repetition.ts:31-34


def test_twist_kat(oracle):
    """twist.ts:22-33 on the y = 0 plane is the identity (cos 0 = 1, sin 0 = 0)."""
    tor = sm.create_torus(0, 0, 0, 1.3)
    s, _ = _oracle_from(oracle, [sm.create_twist(tor, 3)])
    s0, _ = _oracle_from(oracle, [tor])
    assert s.distance([0.4, 0.0, -0.9])[0] == s0.distance([0.4, 0.0, -0.9])[0]
    # off the plane: rotate by k*y by hand in JS arithmetic
    p = np.array([0.4, 0.25, -0.9], np.float32)
    c, sn = math.cos(3 * float(p[1])), math.sin(3 * float(p[1]))
    tw = [f32(c * float(p[0]) - sn * float(p[2])), float(p[1]), f32(sn * float(p[0]) + c * float(p[2]))]
    assert s.distance(p)[0] == s0.distance(tw)[0]


def test_animated_translate_kat(oracle):
    """animatedTranslate.ts:34-48: the child is evaluated at p - dir * sin(time*speed) * amplitude."""
    sph = sm.create_sphere(0, 0, 0, 1.0)
    s, _ = _oracle_from(oracle, [sm.create_animated_translate(sph, (2, 0, 0), 3.0, 0.005)])  # direction is normalised
    assert s.distance([5.0, 0.0, 0.0])[0] == 4.0       # time 0: no offset
    s.set_time(100.0)
    off = f32(math.sin(100.0 * 0.005) * 3.0)
    assert s.distance([5.0, 0.0, 0.0])[0] == f32(5.0 - off) - 1.0
    assert s.object_geometry(0)[1] == 1.0 + 3.0        # bounding radius grows by the amplitude


def test_smooth_union_geometry_kat(oracle):
    """smoothUnion.ts:37-59: midpoint of the centres, max radius + half the centre distance."""
    s, _ = _oracle_from(oracle, [sm.create_smooth_union(sm.create_sphere(-1, 0, 0, 0.5), sm.create_sphere(3, 0, 0, 0.25), 0.1)])
    w, r, bmin, bmax = s.object_geometry(0)
    assert list(w) == [1.0, 0.0, 0.0] and r == 0.5 + 2.0
    assert list(bmin) == [f32(1 - 2.5 * 1.5), f32(-2.5 * 1.5), f32(-2.5 * 1.5)]


# ------------------------------------------------------------------------------------------ builders over trees
@pytest.mark.parametrize("idx", sm.OPERATOR_PRESETS)
def test_native_builders_over_operator_trees_match_oracle(oracle, idx):
    pl = sm.get_preset(idx)
    t, m, q = pl.arrays()
    o = oracle.OracleScene().load_preset(idx)
    o.build_accel("BVH")
    nodes, nn, leaf = rb.build_bvh_scene(t, m, q, pl.op_nodes, pl.object_root)
    b, links, lf = o.bvh_flat()
    arr = np.frombuffer(nodes, dtype=np.dtype([("b", np.float32, 6), ("l", np.int32, 4)]), count=nn)
    assert nn == len(b)
    assert np.array_equal(arr["b"].view(np.uint32), b.view(np.uint32))
    assert np.array_equal(arr["l"], links) and np.array_equal(leaf, lf)
    o.build_accel("Octree")
    nodes, nn, leaf = rb.build_octree_scene(t, m, q, pl.op_nodes, pl.object_root)
    ob, ol, lev, emp, mind, olf = o.octree_flat()
    assert nn == len(ob) and np.array_equal(leaf, olf)
    assert np.array_equal(np.array([nodes[k].min_distance for k in range(nn)]), mind)
    assert np.array_equal(np.array([nodes[k].is_empty for k in range(nn)], np.uint8), emp)
    got_b = np.array([[*nodes[k].bmin, *nodes[k].bmax] for k in range(nn)], np.float32)
    assert np.array_equal(got_b.view(np.uint32), ob.view(np.uint32))


def test_scene_builders_equal_flat_builders_on_plain_lists():
    t, m, q = sm.get_preset(3).arrays()
    a, na, la = rb.build_bvh(t, m, q)
    b, nb, lb = rb.build_bvh_scene(t, m, q)
    assert na == nb and bytes(a) == bytes(b) and np.array_equal(la, lb)


# ------------------------------------------------------------------------------------------ validation of node arrays
def _build_rc(types, w2l, params, nodes, roots):
    L = _lib.lib()
    s, keep = rb.renderer._fill_scene(types, w2l, params, nodes, roots)
    nn, nl = C.c_int32(0), C.c_int32(0)
    return L.rm_build_bvh_scene(C.byref(s), 0, None, C.byref(nn), None, C.byref(nl))


def test_tree_validation_rejects_bad_structures():
    pl = sm.get_preset(16)  # Round(Twist(Box))
    t, m, q = pl.arrays()
    assert _build_rc(t, m, q, pl.op_nodes, pl.object_root) == 0
    bad = pl.op_nodes.copy()
    bad["child"][1][0] = 0  # twist -> round: a cycle
    assert _build_rc(t, m, q, bad, pl.object_root) == _lib.RM_ERR_ARG
    bad = pl.op_nodes.copy()
    bad["child"][0][0] = 99
    assert _build_rc(t, m, q, bad, pl.object_root) == _lib.RM_ERR_ARG
    bad = pl.op_nodes.copy()
    bad["prim"][2] = 5
    assert _build_rc(t, m, q, bad, pl.object_root) == _lib.RM_ERR_ARG
    bad = pl.op_nodes.copy()
    bad["kind"][0] = 9  # e.g. a Mandelbulb or an operator this library does not know
    assert _build_rc(t, m, q, bad, pl.object_root) == _lib.RM_ERR_UNSUPPORTED_PRIMITIVE
    assert _build_rc(t, m, q, pl.op_nodes, np.array([7], np.int32)) == _lib.RM_ERR_ARG
    # nesting deeper than RM_MAX_TREE_DEPTH
    node = sm.create_sphere(0, 0, 0, 1)
    for _ in range(_lib.RM_MAX_TREE_DEPTH):
        node = sm.create_round(node, 0.01)
    deep = sm.flatten([node])
    assert _build_rc(*deep.arrays(), deep.op_nodes, deep.object_root) == _lib.RM_ERR_ARG
    ok = sm.flatten([node.children[0]])
    assert _build_rc(*ok.arrays(), ok.op_nodes, ok.object_root) == 0
