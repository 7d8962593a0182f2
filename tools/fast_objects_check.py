import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import cpu_raymarcher_b200 as rb
W, H = 1920, 1080
for fo in ("1", "0"):
    os.environ["RM_FAST_OBJECTS"] = fo
    w = rb.RaymarchWorker(0)
    for preset in (13, 17):
        job = dict(width=W, height=H, time=0.0, yStart=0, yEnd=H, camera=dict(pitch=0.0, yaw=0.0), algorithm="sphere-tracer",
                   scenePresetIndex=preset, accelerationStructure="None", overshootFactor=1.2, stepSize=0.1)
        w.on_message(job)
        f = w.on_message(job)
        st = w.stats()
        print("fast_objects", fo, "preset", preset, "kernel ms %.2f" % st["kernel_ms"], "sum_iters", st["sum_iters_full"], "sum_sdf", st["sum_sdf_full"], "hits", st["n_hit"], flush=True)
    w.close()
