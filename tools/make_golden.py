"""Generate tests/golden/*.npz from the oracle (the reference itself cannot run here: TypeScript, no JS engine,
gl-matrix not vendored).  Small frames so the fixtures stay small; every algorithm x accel combination."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
cases = []


def add(name, preset, accel, alg, W=48, H=32, pitch=0.0, yaw=0.0, step=0.1, over=1.2, synthetic=None, time=0.0):
    cases.append(dict(name=name, preset=preset, accel=accel, alg=alg, W=W, H=H, pitch=pitch, yaw=yaw, step=step, over=over,
                      synthetic=synthetic, time=time))


add("cfg1_sphere_none_st", 0, "None", "sphere-tracer", 64, 64)
add("cfg2_grid_bvh_st", 2, "BVH", "sphere-tracer", 96, 54)
for alg in ("fixed-step", "adaptive-step", "sphere-tracer"):
    add(f"cfg3_dense_octree_{alg}", 3, "Octree", alg, 96, 54)
add("cfg4_synth2000_bvh_st", 1, "BVH", "sphere-tracer", 48, 27, synthetic=[2000, 0x5EED0001])
for p in (4, 5, 7, 8, 9):
    add(f"cfg5_preset{p}_none_yaw", p, "None", "sphere-tracer", 64, 36, yaw=0.015 * 40)
add("torus_bvh_v2", 5, "BVH", "adaptive-step-v2", yaw=1.0, pitch=0.3)
add("boxes_octree_v3", 9, "Octree", "adaptive-step-v3", yaw=2.0, pitch=-0.4, over=1.7)
add("atom_bvh_fixed_small_step", 4, "BVH", "fixed-step", step=0.03)
add("random_octree_adaptive", 1, "Octree", "adaptive-step", yaw=4.0)
add("spherecube_bvh_v3", 8, "BVH", "adaptive-step-v3", yaw=0.7)
# SDF operator presets (src/util/primitive_operations): one per operator scene, all three structures and several algorithms
add("op_roundedbox_none_st", 6, "None", "sphere-tracer", yaw=0.4, pitch=0.2)
add("op_smoothunion_bvh_v2", 10, "BVH", "adaptive-step-v2", yaw=1.1)
add("op_smoothsub_octree_st", 11, "Octree", "sphere-tracer", yaw=0.6, pitch=0.3)
add("op_animated_bvh_st_t314", 12, "BVH", "sphere-tracer", time=314.0)
add("op_twistedtorus_none_fixed", 14, "None", "fixed-step", yaw=0.3)
add("op_infinitespheres_bvh_st", 15, "BVH", "sphere-tracer", yaw=0.2, pitch=0.1)
add("op_screw_octree_v3", 16, "Octree", "adaptive-step-v3", yaw=2.2)
add("op_chicken_bvh_st", 17, "BVH", "sphere-tracer", yaw=0.8, pitch=0.2)
add("op_67_none_adaptive", 18, "None", "adaptive-step", yaw=0.1)

for c in cases:
    s = po.OracleScene()
    if c["synthetic"]:
        s.load_synthetic(*c["synthetic"])
    else:
        s.load_preset(c["preset"])
    s.build_accel(c["accel"]).set_camera(c["pitch"], c["yaw"]).set_time(c["time"])
    f = s.render(c["W"], c["H"], c["alg"], step_size=c["step"], overshoot=c["over"])
    np.savez_compressed(os.path.join(OUT, c["name"] + ".npz"), depth=f.depth, normal=f.normal, sdfEval=f.sdfEval, iters=f.iters,
                        depth_f64=f.depth_f64, sdf_full=f.sdf_full,
                        phong=po.shade("phong", f.depth, f.normal, f.sdfEval, f.iters, c["W"], c["H"]),
                        sdf_heat=po.shade("sdf-heatmap", f.depth, f.normal, f.sdfEval, f.iters, c["W"], c["H"]),
                        iter_heat=po.shade("iteration-heatmap", f.depth, f.normal, f.sdfEval, f.iters, c["W"], c["H"]))
with open(os.path.join(OUT, "manifest.json"), "w") as fh:
    json.dump({"generator": "tools/make_golden.py (oracle/liboracle.so)", "cases": cases}, fh, indent=1)
print(len(cases), "cases;", sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)) // 1024, "KiB")
