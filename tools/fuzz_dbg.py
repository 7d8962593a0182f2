"""Debug helper: one fuzz seed of tests/test_gpu_operators.py::test_validation_random_operator_trees, pixel by pixel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import cpu_raymarcher_b200 as rb
from cpu_raymarcher_b200 import scene_manager as sm
from cpu_raymarcher_b200.camera import Camera
from oracle import pyoracle as po
import test_gpu_operators as T

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 5
rng = np.random.default_rng(1000 + seed)
has_twist = []
objs = [T._random_tree(rng, int(rng.integers(1, 4)), has_twist) for _ in range(int(rng.integers(1, 4)))]
accel = ["None", "Octree", "BVH"][seed % 3]
alg = T.ALGS[seed % len(T.ALGS)]
pitch, yaw, time = float(rng.uniform(-0.8, 0.8)), float(rng.uniform(0, 6.28)), float(rng.uniform(0, 500))
W, H = 72, 40
pl = sm.flatten(objs)
t, m, q = pl.arrays()
print("accel", accel, "alg", alg, "objects", pl.n_objects, "time", time)
for a in (accel, "None"):
    osc = po.OracleScene()
    osc.set_tree(t, m, q, pl.op_nodes, pl.object_root) if pl.op_nodes is not None else osc.set_prims(t, m, q)
    ref = osc.build_accel(a).set_camera(pitch, yaw).set_time(time).render(W, H, alg)
    ctx = rb.Context(0, validate_fp64=True)
    ctx.upload_scene(t, m, q, a, op_nodes=pl.op_nodes, object_root=pl.object_root)
    cam = Camera()
    cam.set_angles(pitch, yaw)
    f = ctx.render(rb.Context.make_request(W, H, cam.get_rotation_matrix3(), cam.get_position(), algorithm=alg, time=time), extras=True)
    ctx.close()
    bad = np.nonzero((f.sdfEval != ref.sdfEval) | (f.depth != ref.depth))[0]
    print(a, "mismatching pixels", len(bad), "of", W * H)
    for i in bad[:5]:
        print("  px", i, "gpu sdf/iters/depth", f.sdfEval[i], f.iters[i], f.depth_f64[i], "| oracle", ref.sdfEval[i], ref.iters[i], ref.depth_f64[i])
    if a == "BVH":
        b, l, lf = osc.bvh_flat()
        print("  bvh nodes", b.tolist(), l.tolist(), lf.tolist())
        for k in range(pl.n_objects):
            print("  object", k, osc.object_geometry(k))
