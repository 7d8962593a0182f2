"""Kernel time of config 4 when one GPU renders only every N-th 8-row stripe (what each rank of an N-GPU run does):
isolates the end-of-frame tail from the multi-GPU plumbing.  python tools/stripe_time.py (needs a GPU)."""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import cpu_raymarcher_b200 as rb
w = rb.RaymarchWorker(0)
W, H = 3840, 2160
job = dict(width=W, height=H, time=0.0, yStart=0, yEnd=H, camera=dict(pitch=0.0, yaw=0.0), algorithm="sphere-tracer",
           scenePresetIndex=1, accelerationStructure="BVH", overshootFactor=1.2, stepSize=0.1, synthetic=(100000, 0x5EED0001))
sc = w._ensure_scene(1, "BVH", (100000, 0x5EED0001))
for cnt in (1, 2, 4, 8, 16, 32, 64, 135, 270):
    rq = rb.Context.make_request(W, H, sc.camera.get_rotation_matrix3(), sc.camera.get_position(), stripes=(8, cnt, 0), shader="iteration-heatmap")
    ts = []
    for _ in range(4):
        w.ctx.render(rq)
        ts.append(w.ctx.stats()["kernel_ms"])
    print(cnt, "stripe share 1/%d" % cnt, "kernel ms", round(min(ts[1:]), 2), "ideal", round(36.4 / cnt, 2), "passes", w.ctx.stats()["tc_passes"])
