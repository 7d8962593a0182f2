"""Small cases for compute-sanitizer: every kernel variant incl. the CTA-cooperative BVH path (shared request ring,
mbarrier/TMA staging)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpu_raymarcher_b200 as rb
for validate in (False, True):
    w = rb.RaymarchWorker(validate_fp64=validate)
    for preset, syn, accel, alg in ((1, (1500, 0x5EED0001), "BVH", "sphere-tracer"), (3, None, "Octree", "adaptive-step-v3"),
                                    (8, None, "None", "fixed-step"), (2, None, "BVH", "adaptive-step-v2"), (1, (600, 7), "None", "sphere-tracer")):
        f = w.on_message(dict(width=96, height=64, yStart=0, yEnd=64, camera=dict(pitch=0.1, yaw=0.3), algorithm=alg, scenePresetIndex=preset,
                              accelerationStructure=accel, synthetic=syn), shader="phong")
        print(validate, preset, accel, alg, int(f.sdfEval.sum()), w.stats()["kernel_ms"])
    w.close()
