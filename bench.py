#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 raymarch hot path (BASELINE.json metric: Mrays/s & SDF evals/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg4]

One "step" = one frame of the workload.  Default workload = BASELINE.json configs[3]: "Random Spheres"
scaled to 100 000 seeded spheres, sphere tracing + BVH, iteration-heatmap shader, 3840x2160 — the
configuration the north-star roofline target is quoted on (>= 10k primitives, 4K, sphere-traced BVH).
Other configs (--workload cfg1|cfg2|cfg3|cfg5) are available for inspection; they are parity-test
cases, not the bench line.

Prints ONE JSON line (see the contract in the task statement): value = Mrays/s with the scene resident
in HBM and outputs left on the device; e2e = the same metric through the reference-facing worker call
(RaymarchWorker.on_message -> rm_render, host buffers, D2H inside the timed region); roofline = the
render kernel's algorithmic FLOP rate against the FP32 FFMA peak measured live on the same GPU;
cpu_baseline = the oracle (C++ restatement of the reference's TS path) timed on a bounded row sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (description, preset, synthetic_n, accel, algorithm, shader, W, H)
    "cfg1": ("Sphere, sphere tracing, no accel, normal shader, 512x512", 0, None, "None", "sphere-tracer", "normal", 512, 512),
    "cfg2": ("Grid of Spheres, sphere tracing + BVH, Phong, 1920x1080", 2, None, "BVH", "sphere-tracer", "phong", 1920, 1080),
    "cfg3": ("Dense Sphere Grid, sphere tracing + octree, SDF heatmap, 1920x1080", 3, None, "Octree", "sphere-tracer", "sdf-heatmap", 1920, 1080),
    "cfg4": ("Random Spheres scaled to 100k synthetic primitives, sphere tracing + BVH, iteration heatmap, 3840x2160",
             1, 100000, "BVH", "sphere-tracer", "iteration-heatmap", 3840, 2160),
    # search-loop microbenchmark: every scene-distance query evaluates all 100 000 spheres (no acceleration structure)
    "dense100k": ("Random Spheres scaled to 100k synthetic primitives, sphere tracing, NO acceleration structure, normal shader, 384x216",
                  1, 100000, "None", "sphere-tracer", "normal", 384, 216),
    "cfg5": ("Atom, sphere tracing, no accel, normal shader, 7680x4320 (one frame of the analytics sweep)",
             4, None, "None", "sphere-tracer", "normal", 7680, 4320),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--prims", type=int, default=None, help="override the synthetic primitive count (cfg4)")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    return ap.parse_args()


def workload(args):
    desc, preset, syn, accel, alg, shader, W, H = WORKLOADS[args.workload]
    if syn is not None and args.prims:
        syn = args.prims
    W = args.width or W
    H = args.height or H
    return dict(name=args.workload, desc=desc, preset=preset, synthetic=(syn, 0x5EED0001) if syn else None, accel=accel,
                algorithm=alg, shader=shader, W=W, H=H, n_prims=syn)


# ----------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md "clocks DURING the timed region")
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.dev = device_index
        self.samples = []
        self.source = "nvidia-smi, 200 ms period"
        self._stop = threading.Event()
        self._t = None

    def _nvml(self):
        """NVML handle of the CUDA device (what nvidia-smi reads, without forking it: ~10 ms per sample instead of ~200)."""
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(self.dev).uuid)).encode())
        except Exception:
            if os.environ.get("CUDA_VISIBLE_DEVICES"):
                raise
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.dev)

    def _run(self):
        try:
            nv, h = self._nvml()
            reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = ((0x8, 5), (0x40, 6), (0x20, 7), (0x4, 8))  # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap -> column
            self.source = "nvml (the counters nvidia-smi reads), 20 ms period"
            while not self._stop.is_set():
                r = int(reasons(h))
                row = [str(self.dev), str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx), "", hex(r), "", "", "", ""]
                for bit, col in bits:
                    row[col] = "Active" if r & bit else "Not Active"
                self.samples.append(row)
                self._stop.wait(0.02)
            return
        except Exception:
            pass  # no NVML binding / handle: fall back to polling nvidia-smi
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.dev)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[1]))
                mx.append(float(s[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


# ----------------------------------------------------------------------------------------------
# CPU baseline: the oracle on a bounded row sample
# ----------------------------------------------------------------------------------------------
def oracle_scene(wl):
    from oracle import pyoracle as po
    s = po.OracleScene()
    t0 = time.perf_counter()
    if wl["synthetic"]:
        s.load_synthetic(*wl["synthetic"])
    else:
        s.load_preset(wl["preset"])
    t1 = time.perf_counter()
    s.build_accel(wl["accel"])  # the reference rebuilds this per job per frame (raymarchWorker.ts:37-38)
    t2 = time.perf_counter()
    s.set_camera(0.0, 0.0)
    return s, (t1 - t0) * 1e3, (t2 - t1) * 1e3


def cpu_sample_rows(s, wl, target_s: float, threads: int):
    """Pick evenly spaced rows so that one sample takes about target_s seconds."""
    W, H = wl["W"], wl["H"]
    probe = np.linspace(0, H - 1, num=min(H, max(threads, 4)), dtype=np.int32)
    t0 = time.perf_counter()
    s.render_rows(W, H, probe, wl["algorithm"], nthreads=threads)
    per_row = (time.perf_counter() - t0) / len(probe)
    n = int(max(threads, min(H, target_s / max(per_row, 1e-9))))
    n = max(threads, (n // threads) * threads)
    n = min(n, H)
    return np.unique(np.linspace(0, H - 1, num=n, dtype=np.int32))


def run_cpu(s, wl, rows, threads):
    W, H = wl["W"], wl["H"]
    t0 = time.perf_counter()
    f = s.render_rows(W, H, rows, wl["algorithm"], nthreads=threads)
    dt = time.perf_counter() - t0
    rays = len(rows) * W
    return dict(seconds=dt, rays=rays, mrays_s=rays / dt / 1e6, evals_s=float(f.sdf_full.astype(np.uint64).sum()) / dt, frame=f)


def parity_block(wl, job, rows, ref, device: int) -> dict:
    """The oracle rows rendered for cpu_baseline, compared with the same rows of the GPU frame (checker leg, untimed):
    the fast build's full frame against the north-star bar, and the fp64 validation build on exactly those rows (1-row
    band requests) bit for bit."""
    import cpu_raymarcher_b200 as rb
    from oracle import compare as cmp
    from oracle import pyoracle as po
    W, H = wl["W"], wl["H"]
    w = rb.RaymarchWorker(device=device)
    f = w.on_message(job, shader=wl["shader"], extras=True)
    g = cmp.take_rows(f, W, rows)
    want = po.shade(wl["shader"], ref.depth, ref.normal, ref.sdfEval, ref.iters, W, len(rows))
    fast = cmp.fast_agreement(g, ref, want)
    w.close()
    del f
    v = rb.RaymarchWorker(device=device, validate_fp64=True)
    parts = [v.on_message(dict(job, yStart=int(y), yEnd=int(y) + 1), extras=True) for y in rows]
    v.close()

    class G:
        pass

    gv = G()
    for k in ("depth", "normal", "sdfEval", "iters", "depth_f64", "sdf_u32"):
        setattr(gv, k, np.concatenate([getattr(b, k) for b in parts]))
    exact = cmp.bit_exact(gv, ref)
    return {"rows": int(len(rows)), "pixels": fast["pixels"], "px_agree": fast["px_agree"], "hit_agree": fast["hit_agree"],
            "rgb_max": fast["rgb_max"], "rgb_within_1": fast["rgb_within_1"], "depth_rel_max": fast["depth_rel_max"],
            "depth_within_1e-4": fast["depth_within_1e-4"], "counters_equal": fast["counters_equal"],
            "bar": "fast build: >= 99.9 % of pixels agree on hit mask, RGB within 1/255 (normal plane + the workload's shader), depth rel. err <= 1e-4",
            "validation_fp64": dict(exact, note="fp64 validation build on the same rows, compared bit for bit (counters, hit mask, depth bits, bytes)"),
            "against": "oracle (C++ restatement of the reference TS; parity unpinned to the reference itself, see DESIGN.md)",
            "pass": bool(fast["pass"] and exact["all"])}


def reference_arm(args, wl, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the TS original cannot run
    here: no JS engine, gl-matrix not vendored) on all host threads, each step a bounded sample of the workload."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    s, t_scene, t_accel = oracle_scene(wl)
    rows = cpu_sample_rows(s, wl, target_s=8.0, threads=threads)
    for _ in range(min(args.warmup, 1)):
        run_cpu(s, wl, rows[: max(threads, len(rows) // 4)], threads)
    t_all, rays, evals = 0.0, 0, 0.0
    for _ in range(args.steps):
        r = run_cpu(s, wl, rows, threads)
        t_all += r["seconds"]
        rays += r["rays"]
        evals += r["evals_s"] * r["seconds"]
    value = rays / t_all / 1e6
    sample = f"{len(rows)} of {wl['H']} rows (evenly spaced) x {wl['W']} px per step; scene+accel prebuilt once ({t_accel:.0f} ms accel build not timed)"
    line = {"impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_all / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "n_prims": wl["n_prims"], "width": wl["W"], "height": wl["H"]},
            "sdf_evals_per_s": evals / t_all,
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample,
                             "accel_build_ms": t_accel},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# the B200 arm
# ----------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    wl = workload(args)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, wl, rank)
        return

    import torch
    import torch.distributed as dist

    import cpu_raymarcher_b200 as rb
    from cpu_raymarcher_b200 import _lib
    from cpu_raymarcher_b200 import multigpu

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the raymarch path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W, H = wl["W"], wl["H"]

    worker = rb.RaymarchWorker(device=local_rank)
    ctx = worker.ctx
    job = dict(width=W, height=H, time=0.0, yStart=0, yEnd=H, camera=dict(pitch=0.0, yaw=0.0), algorithm=wl["algorithm"],
               scenePresetIndex=wl["preset"], accelerationStructure=wl["accel"], overshootFactor=1.2, stepSize=0.1,
               synthetic=wl["synthetic"])
    # scene: built on rank 0, broadcast over NCCL, uploaded once and kept resident in HBM
    t0 = time.perf_counter()
    sharder = multigpu.FrameSharder(worker, rank, world, local_rank)
    sharder.setup_scene(job)
    upload_ms = (time.perf_counter() - t0) * 1e3
    peak_tflops = ctx.probe_fp32_peak()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    sampler = ClockSampler(local_rank)

    def step(timed: bool):
        flush.zero_()  # L2 flush between iterations (not timed)
        torch.cuda.synchronize()
        st = sharder.render_frame(job, shader=wl["shader"])  # rm_render_device on every rank + gather + stats allreduce
        return st

    for _ in range(args.warmup):
        step(False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    t_dev_ms, kern_ms, flops, xflops, evals, launches = 0.0, 0.0, 0.0, 0.0, 0, 0
    wall0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        st = step(True)
        t_dev_ms += st["frame_ms"]        # max over ranks of the device time of the step (kernel + fused gather)
        kern_ms += st["kernel_ms_max"]
        flops += st["algorithmic_flops"]
        xflops += st["executed_flops"]
        evals += st["sum_sdf_full"]
        launches += st["n_launches"]
        last = st
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    clocks = sampler.stop()

    rays = W * H * args.steps
    value = rays / (t_dev_ms * 1e-3) / 1e6
    achieved_tf = flops / (kern_ms * 1e-3) / 1e12 / world  # per-GPU rate of the dominant kernel
    out = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "n_prims": last["n_prims"], "width": W, "height": H, "accel": wl["accel"],
                   "algorithm": wl["algorithm"], "shader": wl["shader"], "l2": "flushed between steps (256 MiB write)",
                   "parallelism": f"row-stripe x{world}" if world > 1 else "single GPU",
                   "field_math": "fp32 (SDF evaluations)", "control_math": "fp64, unfused (JS-exact ray/interval logic)"},
        "sdf_evals_per_s": evals / (t_dev_ms * 1e-3),
        "avg_sdf_calls_per_pixel": evals / rays, "hit_fraction": last["n_hit"] / (W * H),
        "gpu_launches": launches, "wall_ms_per_step": wall_ms / args.steps, "scene_upload_ms": upload_ms,
        "clocks": clocks,
        "roofline": {"bound": "fp32", "achieved": achieved_tf, "peak": peak_tflops, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tflops if peak_tflops else None, "traffic": None,
                     "peak_source": "measured live: rm_probe_fp32_peak (independent FFMA chains, all SMs, burst)",
                     "flops_per_eval": "reference-equivalent: every SDF call the reference counts x translation-only sphere 7 (screened "
                                       "search, >= 512 spheres) / 11 | general sphere 26 / box 38 / torus 29",
                     "executed_tflops": xflops / (kern_ms * 1e-3) / 1e12 / world,
                     "frac_executed": (xflops / (kern_ms * 1e-3) / 1e12 / world) / peak_tflops if peak_tflops else None,
                     "note": ("achieved / frac = the north star's figure: SDF evals/s as the reference counts them x FLOPs per evaluation / FP32 peak "
                              "(reference-equivalent work); frac_executed = FLOPs this kernel really issued / peak. With translation-only spheres behind a BVH "
                              "the all-primitives fallback is answered exactly by a tensor-core cluster screen, so far fewer "
                              "FLOPs are executed than the reference's brute force implies (executed_tflops); the kernel is then "
                              "bound by the divergent ray/BVH control path, not by the FP32 pipe") if last.get("tc_passes") else None,
                     "tc": {"passes": last.get("tc_passes", 0), "requests": last.get("tc_requests", 0), "items": last.get("tc_items", 0)},
                     "kernel_ms_per_step": kern_ms / args.steps},
    }

    # DRAM traffic of the dominant kernel from the committed ncu --set full capture of this exact workload (else null)
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic_cfg4.json")) as fh:
            tr = json.load(fh)
        if args.workload == tr["workload"] and last["n_prims"] == tr["n_prims"] and (W, H) == (tr["width"], tr["height"]) and world == 1:
            out["roofline"]["traffic"] = tr["traffic_bytes"]
            out["roofline"]["traffic_source"] = tr["source"]
    except Exception:
        pass

    # ---- e2e: through the reference-facing worker call with host buffers (rank 0 band set, N ranks in parallel)
    if not args.no_e2e:
        e2e = sharder.e2e_frames(job, wl["shader"], steps=max(1, min(args.steps, 3)))
        out["e2e"] = {"value": W * H / (e2e["ms_per_frame"] * 1e-3) / 1e6, "unit": "Mrays/s",
                      "h2d_bytes_per_step": e2e["h2d_bytes"], "d2h_bytes_per_step": e2e["d2h_bytes"],
                      "ms_per_step": e2e["ms_per_frame"], "path": e2e["path"]}

    # ---- CPU baseline (rank 0, N=1 only): oracle on a bounded sample of the same workload
    parity_failed = False
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        s, t_scene, t_accel = oracle_scene(wl)
        rows = cpu_sample_rows(s, wl, target_s=12.0, threads=threads)
        r = run_cpu(s, wl, rows, threads)
        out["cpu_baseline"] = {"value": r["mrays_s"], "unit": "Mrays/s", "cores": threads, "kind": "port",
                               "sample": f"{len(rows)} of {H} rows (evenly spaced) x {W} px, {r['seconds']:.1f} s; "
                                         "C++ restatement of the reference TS path (not V8)",
                               "sdf_evals_per_s": r["evals_s"], "accel_build_ms": t_accel}
        if threads > 4:
            # the reference never runs more than 4 workers (main.ts:318): the same sample at its real cap
            sub = rows[np.linspace(0, len(rows) - 1, num=min(len(rows), max(4, len(rows) // (threads // 4))), dtype=np.int64)]
            r4 = run_cpu(s, wl, np.unique(sub), 4)
            out["cpu_baseline"]["value_at_4_workers"] = r4["mrays_s"]
            out["cpu_baseline"]["sample_at_4_workers"] = f"{len(np.unique(sub))} rows, {r4['seconds']:.1f} s (the reference's worker cap, main.ts:318)"
        # ---- parity on the workload the numbers are quoted on: those oracle rows against the GPU frame
        if not args.no_parity:
            out["parity"] = parity_block(wl, job, rows, r["frame"], local_rank)
            parity_failed = not out["parity"]["pass"]
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if parity_failed:
        raise SystemExit("bench.py: the GPU frame does not meet the parity bar against the oracle rows (see \"parity\" in the line above)")


if __name__ == "__main__":
    main()
