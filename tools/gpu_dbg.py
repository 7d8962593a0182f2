import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pyoracle as po
import cpu_raymarcher_b200 as rb
print("cpus", os.cpu_count())
fast = rb.RaymarchWorker(validate_fp64=False)
for W,H in ((160,90),(512,288),(512,512)):
  for preset, accel, alg in ((0,"None","sphere-tracer"),(0,"None","adaptive-step-v2"),(0,"Octree","adaptive-step-v3"),(5,"None","sphere-tracer"),(9,"None","sphere-tracer"),(1,"BVH","sphere-tracer"),(3,"None","sphere-tracer")):
    t0=time.time()
    ref = po.OracleScene().load_preset(preset).build_accel(accel).set_camera(0.1,0.4).render(W,H,alg)
    to=time.time()-t0
    job = dict(width=W, height=H, yStart=0, yEnd=H, camera=dict(pitch=0.1,yaw=0.4), algorithm=alg, scenePresetIndex=preset, accelerationStructure=accel)
    f = fast.on_message(job, shader="phong", extras=True)
    hit_ref, hit_f = ref.depth_f64 < 10, f.depth_f64 < 10
    c_hit = hit_ref != hit_f
    c_n = np.abs(f.normal.reshape(-1,3).astype(int)-ref.normal.reshape(-1,3).astype(int)).max(1) > 1
    want = po.shade("phong", ref.depth, ref.normal, ref.sdfEval, ref.iters, W, H).reshape(-1,4)
    c_rgb = np.abs(f.rgba.reshape(-1,4).astype(int)-want.astype(int)).max(1) > 1
    rel = np.abs(f.depth_f64-ref.depth_f64)/np.abs(ref.depth_f64)
    c_d = rel > 1e-4
    bad = c_hit|c_n|c_rgb|c_d
    print(f"{W}x{H} p{preset} {accel} {alg}: ok={1-bad.mean():.5f} hitmask {c_hit.sum()} normal {c_n.sum()} rgb {c_rgb.sum()} depth {c_d.sum()} (hit {(c_d&hit_ref).sum()}, miss {(c_d&~hit_ref).sum()}) oracle {to:.2f}s")
