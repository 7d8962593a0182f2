"""Kernel time of every reference preset (all 19) at a given resolution: python tools/time_presets.py [W H] (needs a GPU)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import cpu_raymarcher_b200 as rb
from cpu_raymarcher_b200 import scene_manager as sm

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
w = rb.RaymarchWorker(0)
print(f"| preset | accel | kernel ms @ {W}x{H} | Mrays/s |\n|---|---|---|---|")
for p in range(sm.get_preset_count()):
    for accel in ("None", "BVH"):
        job = dict(width=W, height=H, time=0.0, yStart=0, yEnd=H, camera=dict(pitch=0.0, yaw=0.0), algorithm="sphere-tracer",
                   scenePresetIndex=p, accelerationStructure=accel, overshootFactor=1.2, stepSize=0.1)
        w.on_message(job)
        best = min(w.on_message(job) and w.stats()["kernel_ms"] for _ in range(3))
        print(f"| {p} {sm.PRESET_NAMES[p]} | {accel} | {best:.3f} | {W * H / best / 1e3:.0f} |")
