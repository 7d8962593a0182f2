"""Run under torchrun with N >= 2 GPUs: every rank renders its row stripes straight into rank 0's frame (CUDA-IPC
fused gather) and, a second time, into one shared page-locked host frame; rank 0 compares the assembled frames and the all-reduced diagnostics with its own single-GPU render."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpu_raymarcher_b200 as rb  # noqa: E402
from cpu_raymarcher_b200 import multigpu  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for validate in (False, True):
    for job in (dict(width=640, height=356, scenePresetIndex=3, accelerationStructure="Octree", algorithm="adaptive-step-v3"),
                dict(width=512, height=300, scenePresetIndex=1, accelerationStructure="BVH", algorithm="sphere-tracer", synthetic=(3000, 0x5EED0001)),
                dict(width=333, height=97, scenePresetIndex=8, accelerationStructure="None", algorithm="fixed-step"),
                dict(width=2048, height=1100, scenePresetIndex=2, accelerationStructure="BVH", algorithm="sphere-tracer"),
                # operator trees travel in the scene broadcast too (presets 11 SmoothSubtraction, 12 SmoothUnion [animated], 16 Screw)
                dict(width=320, height=184, scenePresetIndex=11, accelerationStructure="BVH", algorithm="sphere-tracer", time=1234.5),
                dict(width=320, height=184, scenePresetIndex=12, accelerationStructure="Octree", algorithm="adaptive-step-v2", time=777.0),
                dict(width=320, height=184, scenePresetIndex=16, accelerationStructure="None", algorithm="sphere-tracer")):
        job = dict(dict(time=0.0), **job)
        job = dict(job, yStart=0, yEnd=job["height"], camera=dict(pitch=0.1, yaw=0.5), overshootFactor=1.2, stepSize=0.1)
        w = rb.RaymarchWorker(device=local, validate_fp64=validate)
        sh = multigpu.FrameSharder(w, rank, world, local)
        sh.setup_scene(job)
        st = sh.render_frame(job, shader="phong")
        # the same frame into the shared page-locked HOST frame: every rank downloads its own stripes (twice: reuse)
        for _ in range(2):
            st_h, host = sh.render_frame_host(job, shader="phong")
        if rank == 0:
            got = sh.download_frame("phong")
            solo = rb.RaymarchWorker(device=local, validate_fp64=validate)
            ref = solo.on_message(job, shader="phong")
            rs = solo.stats()
            same = all(np.array_equal(got[k], getattr(ref, k)) for k in ("depth", "normal", "sdfEval", "iters", "rgba"))
            stats_ok = all(st[k] == rs[k] for k in ("n_pixels", "sum_sdf", "sum_iters", "max_sdf", "min_sdf", "sum_sdf_full", "n_hit"))
            same_h = all(np.array_equal(host[k], getattr(ref, k)) for k in ("depth", "normal", "sdfEval", "iters", "rgba"))
            stats_ok = stats_ok and all(st_h[k] == rs[k] for k in ("n_pixels", "sum_sdf", "sum_iters", "max_sdf", "min_sdf", "sum_sdf_full", "n_hit"))
            print(f"world={world} validate={validate} preset={job['scenePresetIndex']} {job['accelerationStructure']}: frame {'OK' if same else 'MISMATCH'}, "
                  f"host frame {'OK' if same_h else 'MISMATCH'}, stats {'OK' if stats_ok else 'MISMATCH'}", flush=True)
            ok = ok and same and same_h and stats_ok
            solo.close()
        dist.barrier()
        sh.release()
        w.close()
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, src=0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
