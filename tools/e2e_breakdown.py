"""Where the end-to-end time of one rm_render goes: kernel (CUDA events), rm_render wall (inside the library), Python wall.
Usage: python tools/e2e_breakdown.py   (cfg4: 100k spheres, BVH, 4K, iteration heatmap; RM_EARLY_COPY=0 switches the early download off)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpu_raymarcher_b200 as rb

w = rb.RaymarchWorker(device=0)
job = dict(width=3840, height=2160, time=0, yStart=0, yEnd=2160, camera=dict(pitch=0.0, yaw=0.0), algorithm="sphere-tracer",
           scenePresetIndex=1, accelerationStructure="BVH", overshootFactor=1.2, stepSize=0.1, synthetic=(100000, 0x5EED0001))
for early in ("1", "0", "1", "0"):
    os.environ["RM_EARLY_COPY"] = early
    w.on_message(job, shader="iteration-heatmap", pinned=True)
    rows = []
    for _ in range(5):
        t0 = time.perf_counter()
        w.on_message(job, shader="iteration-heatmap", pinned=True)
        py = (time.perf_counter() - t0) * 1e3
        st = w.stats()
        rows.append((st["kernel_ms"], st["wall_ms"], py))
    k, wl, py = (sum(c) / len(c) for c in zip(*rows))
    print(f"early={early} kernel {k:.2f} ms  rm_render wall {wl:.2f} ms  python wall {py:.2f} ms", flush=True)
