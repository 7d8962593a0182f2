import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pyoracle as po
import cpu_raymarcher_b200 as rb
W,H=160,90
val = rb.RaymarchWorker(validate_fp64=True)
for preset, accel, alg in ((0,"None","sphere-tracer"),(2,"BVH","adaptive-step")):
    ref = po.OracleScene().load_preset(preset).build_accel(accel).render(W,H,alg)
    job = dict(width=W, height=H, yStart=0, yEnd=H, camera=dict(pitch=0,yaw=0), algorithm=alg, scenePresetIndex=preset, accelerationStructure=accel)
    f = val.on_message(job, extras=True)
    bad = np.nonzero((f.normal.reshape(-1,3) != ref.normal.reshape(-1,3)).any(1))[0]
    print(preset, accel, alg, "n bad", len(bad))
    for i in bad[:8]:
        print("  px", i%W, i//W, "gpu", f.normal.reshape(-1,3)[i], "ref", ref.normal.reshape(-1,3)[i], "depth", repr(ref.depth_f64[i]), float(f.depth_f32[i]), "sdf", f.sdfEval[i], ref.sdfEval[i], "iters", ref.iters[i])
