"""BASELINE config 5: the Analytics page's rotation sweep (main.ts:438-441: yaw += 0.015 before every frame) over
Atom / Torus / Cube / Sphere / Pyramid of Boxes, 360 frames each, with the frame diagnostics of main.ts:527-548
(avg / max / min SDF calls, avg iterations) all-reduced over the ranks.

    python tools/analytics_sweep.py [--width 7680 --height 4320 --frames 360]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/analytics_sweep.py ...

One JSON line per preset (rank 0): frames/s, Mrays/s, and the diagnostics of the last frame."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# BASELINE config 5 / SURVEY.md §8d: presets 4, 5, 7, 8, 9 (sceneManager.ts:159-207).  (Round 1 had preset 0 "Sphere" in place of
# 8 "Sphere and Cube".)  The measured workload is `bench.py --workload cfg5sweep`; this script is the standalone per-preset view.
PRESETS = ((4, "Atom"), (5, "Torus"), (7, "Cube"), (8, "Sphere and Cube"), (9, "Pyramid of Boxes"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=7680)
    ap.add_argument("--height", type=int, default=4320)
    ap.add_argument("--frames", type=int, default=360)
    ap.add_argument("--accel", default="None")
    ap.add_argument("--algorithm", default="sphere-tracer")
    args = ap.parse_args()
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))

    import torch
    import torch.distributed as dist

    import cpu_raymarcher_b200 as rb
    from cpu_raymarcher_b200 import multigpu
    from cpu_raymarcher_b200.camera import Camera

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    worker = rb.RaymarchWorker(device=local_rank)
    sharder = multigpu.FrameSharder(worker, rank, world, local_rank)
    W, H = args.width, args.height
    for preset, name in PRESETS:
        job = dict(width=W, height=H, time=0.0, yStart=0, yEnd=H, camera=dict(pitch=0.0, yaw=0.0), algorithm=args.algorithm,
                   scenePresetIndex=preset, accelerationStructure=args.accel, overshootFactor=1.2, stepSize=0.1)
        sharder.setup_scene(job)
        cam = Camera()  # the controller's camera: the sweep accumulates yaw exactly like main.ts
        sharder.render_frame(job, shader="sdf-heatmap")  # warm-up (allocates the frame planes)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        dev_ms = 0.0
        st = None
        for _ in range(args.frames):
            cam.rotate_camera(0.0, 0.015)
            job["camera"] = dict(pitch=cam.pitch, yaw=cam.yaw)
            st = sharder.render_frame(job, shader="sdf-heatmap")
            dev_ms += st["frame_ms"]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        wall = time.perf_counter() - t0
        if rank == 0:
            n = st["n_pixels"]
            print(json.dumps({
                "preset": name, "frames": args.frames, "width": W, "height": H, "n_gpus": world,
                "frames_per_s_wall": args.frames / wall, "Mrays_per_s_device": W * H * args.frames / (dev_ms * 1e-3) / 1e6,
                "last_frame": {"yaw": cam.yaw, "avg_sdf_calls": st["sum_sdf"] / n, "max_sdf_calls": st["max_sdf"],
                               "min_sdf_calls": st["min_sdf"], "avg_iterations": st["sum_iters"] / n}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
