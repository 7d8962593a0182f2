"""Ad-hoc GPU parity sweep (development aid): validation build vs oracle, bit-exact; fast build, tolerance."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pyoracle as po
import cpu_raymarcher_b200 as rb

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (160, 90)
val = rb.RaymarchWorker(validate_fp64=True)
fast = rb.RaymarchWorker(validate_fp64=False)
cases = []
for preset in (0, 1, 2, 3, 4, 5, 7, 8, 9):
    for accel in ("None", "Octree", "BVH"):
        for alg in ("sphere-tracer", "fixed-step", "adaptive-step", "adaptive-step-v2", "adaptive-step-v3"):
            cases.append((preset, None, accel, alg, 0.0, 0.0))
cases += [(1, (2000, 0x5EED0001), "BVH", "sphere-tracer", 0.0, 0.0), (1, (2000, 0x5EED0001), "Octree", "sphere-tracer", 0.0, 0.0),
          (3, None, "Octree", "sphere-tracer", 0.2, 0.7), (2, None, "BVH", "sphere-tracer", -0.4, 2.5),
          (9, None, "BVH", "adaptive-step-v3", 0.5, 1.0), (5, None, "None", "sphere-tracer", 0.3, 0.3)]
bad = 0
for preset, syn, accel, alg, pitch, yaw in cases:
    osc = po.OracleScene()
    if syn: osc.load_synthetic(*syn)
    else: osc.load_preset(preset)
    osc.build_accel(accel).set_camera(pitch, yaw)
    ref = osc.render(W, H, alg)
    job = dict(width=W, height=H, time=0, yStart=0, yEnd=H, camera=dict(pitch=pitch, yaw=yaw), algorithm=alg,
               scenePresetIndex=preset, accelerationStructure=accel, overshootFactor=1.2, stepSize=0.1, synthetic=syn)
    t0 = time.time(); fv = val.on_message(job, shader="phong", shader_analytics="sdf-heatmap", extras=True); tv = time.time() - t0
    kv = val.stats()["kernel_ms"]
    ex = dict(depth=np.array_equal(fv.depth, ref.depth), normal=np.array_equal(fv.normal, ref.normal),
              sdf=np.array_equal(fv.sdfEval, ref.sdfEval), iters=np.array_equal(fv.iters, ref.iters),
              sdf32=np.array_equal(fv.sdf_u32, ref.sdf_full),
              depth64=np.array_equal(fv.depth_f64.view(np.uint64), ref.depth_f64.view(np.uint64)),
              iters32=np.array_equal(fv.sdf_u32 * 0 + fv.iters, ref.iters),
              phong=np.array_equal(fv.rgba, po.shade("phong", ref.depth, ref.normal, ref.sdfEval, ref.iters, W, H)),
              heat=np.array_equal(fv.rgba_analytics, po.shade("sdf-heatmap", ref.depth, ref.normal, ref.sdfEval, ref.iters, W, H)))
    ff = fast.on_message(job, shader="phong", extras=True)
    kf = fast.stats()["kernel_ms"]
    hit_ref = ref.depth_f64 < 10
    hit_f = ff.depth_f64 < 10
    agree_hit = (hit_ref == hit_f)
    ref_rgb = po.shade("phong", ref.depth, ref.normal, ref.sdfEval, ref.iters, W, H).reshape(-1, 4)[:, :3].astype(int)
    rgb_ok = (np.abs(ff.rgba.reshape(-1, 4)[:, :3].astype(int) - ref_rgb).max(1) <= 1)
    nrm_ok = (np.abs(ff.normal.reshape(-1, 3).astype(int) - ref.normal.reshape(-1, 3).astype(int)).max(1) <= 1)
    both = hit_ref & hit_f
    rel = np.abs(ff.depth_f64[both] - ref.depth_f64[both]) / np.maximum(np.abs(ref.depth_f64[both]), 1e-9) if both.any() else np.zeros(1)
    ok_px = agree_hit & rgb_ok & nrm_ok
    allv = all(ex.values())
    fast_frac = ok_px.mean()
    flag = "" if (allv and fast_frac >= 0.999) else "  <<<<<<"
    if flag: bad += 1
    print(f"p{preset}{'s' if syn else ''} {accel:6s} {alg:16s} val={'OK' if allv else [k for k,v in ex.items() if not v]} ({kv:.2f}ms)  fast: px_ok={fast_frac:.5f} hit_agree={agree_hit.mean():.5f} depth_rel_max={rel.max():.2e} p99.9={np.quantile(rel,0.999):.2e} ({kf:.2f}ms){flag}")
print("BAD CASES:", bad, "of", len(cases))
