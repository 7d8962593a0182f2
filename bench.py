#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 raymarch hot path (BASELINE.json metric: Mrays/s & SDF evals/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg4]

One "step" = one frame of the workload.  Default workload = BASELINE.json configs[3]: "Random Spheres"
scaled to 100 000 seeded spheres, sphere tracing + BVH, iteration-heatmap shader, 3840x2160 — the
configuration the north-star roofline target is quoted on (>= 10k primitives, 4K, sphere-traced BVH).
Other configs (--workload cfg1|cfg2|cfg3|cfg5) are available for inspection; they are parity-test
cases, not the bench line.  --workload cfg5sweep is BASELINE.json configs[4] as specified: the 5-preset, 360-frame 8K
analytics rotation sweep with reduced per-frame diagnostics (one step = the whole sweep; use --steps 1 --warmup 1).
N > 1: under torchrun (the driver's launch) one process per GPU over NCCL; `python bench.py --gpus N` WITHOUT torchrun drives
the N GPUs from this one process through the library's own rm_pool (C ABI).

Prints ONE JSON line (see the contract in the task statement): value = Mrays/s with the scene resident
in HBM and outputs left on the device; e2e = the same metric through the reference-facing worker call
(RaymarchWorker.on_message -> rm_render, host buffers, D2H inside the timed region); roofline = the
FLOPs the render kernel EXECUTED on the FP32 pipe against the FFMA peak measured live on the same GPU;
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
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS) + ["cfg5sweep"])
    ap.add_argument("--frames", type=int, default=360, help="cfg5sweep: frames per preset")
    ap.add_argument("--sweep-mode", default="frames", choices=["frames", "stripes"], help="cfg5sweep at N > 1: frames dealt to GPUs, or every frame split N ways")
    ap.add_argument("--prims", type=int, default=None, help="override the synthetic primitive count (cfg4)")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    return ap.parse_args()


def config_of(wl, n_prims, parallelism):
    """The `config` object of the JSON line — identical keys and values on the b200 arm and on the reference arm."""
    return {"workload": wl["desc"], "n_prims": n_prims, "width": wl["W"], "height": wl["H"], "accel": wl["accel"],
            "algorithm": wl["algorithm"], "shader": wl["shader"], "l2": "flushed between steps (256 MiB write per device)",
            "parallelism": parallelism,
            "field_math": "fp32 (SDF evaluations)", "control_math": "fp64, unfused (JS-exact ray/interval logic)"}


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
            "config": config_of(wl, s.n_prims, "single GPU" if args.gpus <= 1 else f"row-stripe x{args.gpus}"),
            "config_note": "same workload as the b200 arm (its config keys are repeated verbatim); this arm runs the CPU restatement of the "
                           "reference path in f64 on the host cores, on a bounded row sample",
            "sdf_evals_per_s": evals / t_all,
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample,
                             "accel_build_ms": t_accel},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# the B200 arm
# ----------------------------------------------------------------------------------------------
class TorchrunEngine:
    """One process per GPU (N = 1, or N ranks under torchrun): FrameSharder — NCCL scene broadcast, CUDA-IPC fused gather,
    packed stats all-reduces."""

    def __init__(self, rank, world, local_rank):
        import cpu_raymarcher_b200 as rb
        from cpu_raymarcher_b200 import multigpu
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.worker = rb.RaymarchWorker(device=local_rank)
        self.sharder = multigpu.FrameSharder(self.worker, rank, world, local_rank)
        self.mode = "single GPU" if world == 1 else f"row-stripe x{world}, one process per GPU (torchrun + NCCL)"
        self.flush_devices = [local_rank]

    def setup(self, job):
        self.sharder.setup_scene(job)

    def peak(self):
        return self.worker.ctx.probe_fp32_peak()

    def frame(self, job, shader):
        st = self.sharder.render_frame(job, shader=shader)  # rm_render_device on every rank + fused gather + stats all-reduce
        st["kernel_ms_local"] = self.worker.ctx.stats()["kernel_ms"]
        return st

    def per_rank(self, values):
        """all-gather one float per rank (per-rank kernel time: skew vs tail)."""
        if self.world == 1:
            return [values]
        import torch
        import torch.distributed as dist
        t = torch.tensor([values], dtype=torch.float64, device=torch.device("cuda", self.local_rank))
        out = [torch.zeros_like(t) for _ in range(self.world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def e2e(self, job, shader, steps, flush=None):
        return self.sharder.e2e_frames(job, shader, steps=steps, between=flush)

    def n_prims(self):
        return self.worker.ctx.n_prims


class PoolEngine:
    """N GPUs driven by ONE process through the C ABI (rm_pool_*): no torchrun, no process group."""

    def __init__(self, n):
        import cpu_raymarcher_b200 as rb
        self.rank, self.world, self.local_rank = 0, n, 0
        self.pool = rb.RaymarchPool(list(range(n)))
        self.mode = f"row-stripe x{n}, one process (rm_pool: one host thread per GPU, peer-copied scene, host-reduced stats)"
        self.flush_devices = list(range(n))
        self._dev_ms = [0.0] * n

    def setup(self, job):
        self.pool._ensure_scene(int(job.get("scenePresetIndex", 0)), job.get("accelerationStructure", "None"), job.get("synthetic"))

    def peak(self):
        return self.pool.probe_fp32_peak()

    def frame(self, job, shader):
        st = self.pool.render_device(job, shader=shader)  # planes in device 0's HBM, peers store into them (fused gather)
        st["frame_ms"] = st["kernel_ms"]  # slowest device
        st["kernel_ms_max"] = st["kernel_ms"]
        st["n_prims"] = self.pool.n_prims
        for i in range(self.world):
            self._dev_ms[i] += self.pool.device_stats(i)["kernel_ms"]
        st["kernel_ms_local"] = 0.0
        return st

    def per_rank(self, _):
        out, self._dev_ms = self._dev_ms, [0.0] * self.world
        return out

    def e2e(self, job, shader, steps, flush=None):
        import ctypes as C
        from cpu_raymarcher_b200 import _lib
        W, H = int(job["width"]), int(job["height"])
        self.pool.on_message(job, shader=shader, pinned=True)  # warm-up allocates the page-locked planes
        t_all = 0.0
        for _ in range(steps):
            if flush:
                flush()  # L2 flush between frames, not timed
            t0 = time.perf_counter()
            self.pool.on_message(job, shader=shader, pinned=True)
            t_all += time.perf_counter() - t0
        ms = t_all * 1e3 / steps
        return {"ms_per_frame": ms, "h2d_bytes": C.sizeof(_lib.Request) * self.world, "d2h_bytes": W * H * (8 + (4 if shader else 0)),
                "path": "RaymarchPool.on_message -> rm_pool_render: every device downloads its own stripes into the caller's page-locked planes during its render"}

    def n_prims(self):
        return self.pool.n_prims


def ncu_facts(wl_name, n_prims, W, H):
    """Per-launch facts of the dominant kernel from the committed ncu --set full capture of this exact workload (profiles/):
    DRAM traffic, issue-slot utilisation, active lanes per instruction, top stall reasons.  None when no capture matches."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for fn in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if not (fn.startswith("ncu_") and fn.endswith(".json")):
            continue
        try:
            with open(os.path.join(pdir, fn)) as fh:
                tr = json.load(fh)
            if tr["workload"] == wl_name and tr["n_prims"] == n_prims and (tr["width"], tr["height"]) == (W, H):
                best = tr  # the last one by name = the latest round
        except Exception:
            continue
    return best


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload == "cfg5sweep":
        return sweep_main(args, rank, local_rank, world)
    wl = workload(args)
    if args.impl == "reference":
        reference_arm(args, wl, rank)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the raymarch path has no CPU fallback")
    single_process_pool = world == 1 and args.gpus > 1  # `python bench.py --gpus N` without torchrun: the library's own pool
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W, H = wl["W"], wl["H"]
    eng = PoolEngine(args.gpus) if single_process_pool else TorchrunEngine(rank, world, local_rank)
    n_gpus = args.gpus if single_process_pool else world

    job = dict(width=W, height=H, time=0.0, yStart=0, yEnd=H, camera=dict(pitch=0.0, yaw=0.0), algorithm=wl["algorithm"],
               scenePresetIndex=wl["preset"], accelerationStructure=wl["accel"], overshootFactor=1.2, stepSize=0.1,
               synthetic=wl["synthetic"])
    # scene: built once, replicated (NCCL broadcast / peer copies), resident in HBM
    t0 = time.perf_counter()
    eng.setup(job)
    upload_ms = (time.perf_counter() - t0) * 1e3
    peak_tflops = eng.peak()

    flush = [torch.empty(256 << 20, dtype=torch.uint8, device=torch.device("cuda", d)) for d in eng.flush_devices]  # > 126 MB L2
    sampler = ClockSampler(local_rank)

    def flush_l2():
        for f in flush:
            f.zero_()  # L2 flush between iterations (not timed)
        for d in eng.flush_devices:
            torch.cuda.synchronize(d)

    def step():
        flush_l2()
        return eng.frame(job, wl["shader"])

    for _ in range(args.warmup):
        step()
    eng.per_rank(0.0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    t_dev_ms = kern_ms = flops = xflops = fp32flops = tflops_mma = 0.0
    evals = launches = 0
    local_ms = 0.0
    wall0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        st = step()
        t_dev_ms += st["frame_ms"]        # max over ranks / devices of the CUDA-event time of the step's kernel (fused gather inside)
        kern_ms += st["kernel_ms_max"]
        local_ms += st["kernel_ms_local"]
        flops += st["algorithmic_flops"]
        xflops += st["executed_flops"]
        fp32flops += st.get("fp32_pipe_flops", 0.0)
        tflops_mma += st.get("tensor_flops", 0.0)
        evals += st["sum_sdf_full"]
        launches += st["n_launches"]
        last = st
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    clocks = sampler.stop()
    per_rank_ms = [v / args.steps for v in eng.per_rank(local_ms)]

    rays = W * H * args.steps
    value = rays / (t_dev_ms * 1e-3) / 1e6
    ksec = kern_ms * 1e-3
    fp32_tf = fp32flops / ksec / 1e12 / n_gpus       # per-GPU rate of the dominant kernel, FP32 (non-tensor) pipe
    mma_tf = tflops_mma / ksec / 1e12 / n_gpus
    equiv_tf = flops / ksec / 1e12 / n_gpus
    facts = ncu_facts(args.workload, last["n_prims"], W, H) if n_gpus == 1 else None
    out = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_of(wl, last["n_prims"], "single GPU" if n_gpus == 1 else f"row-stripe x{n_gpus}"),
        "route": eng.mode,
        "value_timing": "device-resident: max over GPUs of the CUDA-event kernel time per step (the tile gather is fused into the kernel); "
                        "the stats reduction and host time are in wall_ms_per_step and in e2e, which is the headline",
        "kernel_ms_per_gpu": per_rank_ms,
        "sdf_evals_per_s_reference_counted": evals / (t_dev_ms * 1e-3),
        "avg_sdf_calls_per_pixel": evals / rays, "hit_fraction": last["n_hit"] / (W * H),
        "gpu_launches": launches, "wall_ms_per_step": wall_ms / args.steps, "scene_upload_ms": upload_ms,
        "clocks": clocks,
        "roofline": {
            "bound": "fp32", "achieved": fp32_tf, "peak": peak_tflops, "unit": "TFLOP/s",
            "frac": fp32_tf / peak_tflops if peak_tflops else None,
            "traffic": facts["traffic_bytes"] if facts else None,
            "peak_source": "measured live: rm_probe_fp32_peak (independent FFMA chains, all SMs, burst)",
            "achieved_is": "FLOPs the kernel EXECUTED on the FP32 (non-tensor) pipe for SDF work (rm_stats.fp32_pipe_flops: primitive evaluations actually "
                           "performed x their executed FLOP count — translation-only sphere 11, 7 in the screened search; general sphere 26 / box 38 / "
                           "torus 29) / CUDA-event kernel time / GPUs (SURVEY.md §8d: never more than the executed variant's algorithmic minimum)",
            "tensor_tflops": mma_tf,
            "tensor_note": "tcgen05 tf32 MMAs of the cluster screen (rm_stats.tensor_flops), reported apart: not FP32-pipe work",
            "work_equiv_tflops": equiv_tf,
            "work_equiv_note": "reference-counted SDF calls x FLOPs per call / time: what a brute-force evaluation of the reference's counters would need; "
                               "NOT a roofline fraction (the cluster screen answers the all-primitives fallback without evaluating every sphere)",
            "kernel_ms_per_step": kern_ms / args.steps,
            "tc": {"passes": last.get("tc_passes", 0), "requests": last.get("tc_requests", 0), "items": last.get("tc_items", 0)},
        },
    }
    if facts:
        out["roofline"].update({"traffic_source": facts["source"], "issue_util": facts.get("issue_util"), "lanes_per_inst": facts.get("lanes_per_inst"),
                                "stalls_top": facts.get("stalls_top"),
                                "limiter": "latency, not a pipe: divergent per-ray BVH control path (issue slots mostly empty, few active lanes per instruction); "
                                           "see the ncu fields beside this note"})

    # ---- e2e: through the reference-facing worker call with host buffers
    if not args.no_e2e:
        e2e = eng.e2e(job, wl["shader"], steps=max(3, min(args.steps, 10)), flush=flush_l2)
        out["e2e"] = {"value": W * H / (e2e["ms_per_frame"] * 1e-3) / 1e6, "unit": "Mrays/s",
                      "h2d_bytes_per_step": e2e["h2d_bytes"], "d2h_bytes_per_step": e2e["d2h_bytes"],
                      "ms_per_step": e2e["ms_per_frame"], "path": e2e["path"],
                      "timing": "host wall clock around each call (request in, planes in host memory out), L2 flushed between frames (not timed)"}

    # ---- CPU baseline (rank 0, N=1 only): oracle on a bounded sample of the same workload
    parity_failed = False
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
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


# ----------------------------------------------------------------------------------------------
# BASELINE config 5: the Analytics rotation sweep as a measured workload
# ----------------------------------------------------------------------------------------------
SWEEP_PRESETS = ((4, "Atom"), (5, "Torus"), (7, "Cube"), (8, "Sphere and Cube"), (9, "Pyramid of Boxes"))  # sceneManager.ts:159-207


def sweep_jobs(preset, W, H, frames):
    """The controller's camera (main.ts:438-441): yaw += 0.015 in f64 BEFORE every frame, from 0."""
    from cpu_raymarcher_b200.camera import Camera
    cam = Camera()
    jobs = []
    for _ in range(frames):
        cam.rotate_camera(0.0, 0.015)
        jobs.append(dict(width=W, height=H, time=0.0, yStart=0, yEnd=H, camera=dict(pitch=cam.pitch, yaw=cam.yaw), algorithm="sphere-tracer",
                         scenePresetIndex=preset, accelerationStructure="None", overshootFactor=1.2, stepSize=0.1))
    return jobs


def sweep_main(args, rank, local_rank, world):
    """--workload cfg5sweep: presets 4, 5, 7, 8, 9 x `--frames` (360) frames at 7680x4320, sphere tracing, no acceleration
    structure, normal shader; per-frame diagnostics of main.ts:527-548 reduced over the GPUs.  One step = the whole sweep.
    --sweep-mode frames: different frames on different GPUs (per-frame stats combined with one all-reduce per preset);
    --sweep-mode stripes: every frame split N ways (per-frame all-reduce).  Every 60th frame is parity-sampled against the
    oracle after the timed region."""
    import torch
    import torch.distributed as dist

    import cpu_raymarcher_b200 as rb
    from cpu_raymarcher_b200 import multigpu
    if args.impl == "reference":
        if rank == 0:
            from oracle import pyoracle as po
            threads = os.cpu_count() or 1
            W, H = args.width or 7680, args.height or 4320
            rows = np.arange(0, H, 64, dtype=np.int32)
            t_all, rays = 0.0, 0
            for preset, _ in SWEEP_PRESETS:
                s = po.OracleScene().load_preset(preset).build_accel("None")
                for job in sweep_jobs(preset, W, H, args.frames)[:: max(1, args.frames // 6)]:
                    s.set_camera(job["camera"]["pitch"], job["camera"]["yaw"])
                    t0 = time.perf_counter()
                    s.render_rows(W, H, rows, "sphere-tracer", nthreads=threads)
                    t_all += time.perf_counter() - t0
                    rays += len(rows) * W
            v = rays / t_all / 1e6
            print(json.dumps({"impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0,
                              "ms_per_step": t_all * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                              "config": {"workload": "cfg5sweep", "width": W, "height": H},
                              "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": threads, "kind": "port",
                                               "sample": f"every 64th row of 6 frames per preset ({len(rows)} rows x {W} px each)"},
                              "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the raymarch path has no CPU fallback")
    W, H, frames = args.width or 7680, args.height or 4320, args.frames
    pool_mode = world == 1 and args.gpus > 1
    n_gpus = args.gpus if pool_mode else world
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    pool = rb.RaymarchPool(list(range(n_gpus))) if pool_mode else None
    worker = None if pool_mode else rb.RaymarchWorker(device=local_rank)
    sharder = None if pool_mode else multigpu.FrameSharder(worker, rank, world, local_rank)
    local_sharder = None if pool_mode else multigpu.FrameSharder(worker, 0, 1, local_rank)  # whole frames on this GPU, no collective per frame

    def run_sweep(nframes):
        """-> per preset: list of per-frame diagnostics (identical on every rank), kernel ms summed on the busiest GPU, launches"""
        per_preset, kms, launches = [], 0.0, 0
        for preset, _ in SWEEP_PRESETS:
            jobs = sweep_jobs(preset, W, H, nframes)
            if pool_mode:
                pool._ensure_scene(preset, "None")
                if args.sweep_mode == "frames":
                    sts = pool.render_frames(jobs, shader="normal")
                    tot = pool.stats()
                    kms += tot["kernel_ms"]
                    launches += tot["n_launches"]
                else:
                    sts = []
                    for job in jobs:
                        st = pool.render_device(job, shader="normal")
                        sts.append(st)
                        kms += st["kernel_ms"]
                        launches += st["n_launches"]
            elif args.sweep_mode == "frames" or world == 1:
                worker._ensure_scene(preset, "None")
                mine, my_ms = {}, 0.0
                for k in range(rank, nframes, world):
                    st = local_sharder.render_frame(jobs[k], shader="normal")
                    mine[k] = st
                    my_ms += st["kernel_ms"]
                    launches += st["n_launches"]
                # disjoint rows: one SUM all-reduce per preset carries every frame's diagnostics to every rank
                sts = multigpu.allreduce_frame_table(mine, nframes, dev)
                if world > 1:
                    ms_t = torch.tensor([my_ms], dtype=torch.float64, device=dev)
                    dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
                    my_ms = float(ms_t.item())
                kms += my_ms
            else:
                sharder.setup_scene(jobs[0])
                sts = []
                for job in jobs:
                    st = sharder.render_frame(job, shader="normal")  # per-frame NCCL all-reduce of the diagnostics
                    sts.append(st)
                    kms += st["frame_ms"]
                    launches += st["n_launches"]
            per_preset.append(sts)
        return per_preset, kms, launches

    for _ in range(args.warmup):
        run_sweep(min(frames, 8))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0 = time.perf_counter()
    kms_all, launches_all, last = 0.0, 0, None
    for _ in range(args.steps):
        last, kms, ln = run_sweep(frames)
        kms_all += kms
        launches_all += ln
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    n_frames_total = frames * len(SWEEP_PRESETS) * args.steps
    rays = W * H * n_frames_total
    out = {"metric": "Mrays/s", "value": rays / wall / 1e6, "unit": "Mrays/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": wall * 1e3 / args.steps, "higher_is_better": True, "scaling": "strong" if args.sweep_mode == "stripes" else "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"Analytics rotation sweep over Atom / Torus / Cube / Sphere and Cube / Pyramid of Boxes, {frames} frames each at {W}x{H}, "
                                  "sphere tracing, no acceleration structure, normal shader, per-frame SDF-call / iteration diagnostics reduced over the GPUs",
                      "frames_per_preset": frames, "presets": [p for p, _ in SWEEP_PRESETS], "width": W, "height": H,
                      "parallelism": (("frame-parallel" if args.sweep_mode == "frames" else "row-stripe") + f" x{n_gpus}, " +
                                      ("one process (rm_pool)" if pool_mode else ("one process per GPU (torchrun + NCCL)" if world > 1 else "single GPU"))),
                      "l2": "no flush: every frame differs (rotating camera) and writes 0.4 GB of planes, 3x the 126 MB L2"},
           "value_timing": "wall clock of the whole sweep (barrier + synchronize on both sides), all frames, all GPUs",
           "frames_per_s": n_frames_total / wall, "kernel_ms_busiest_gpu_per_step": kms_all / args.steps,
           "gpu_launches": launches_all, "clocks": clocks,
           "per_preset_last_frame": [{"preset": name, "yaw": sweep_jobs(p, 8, 8, frames)[-1]["camera"]["yaw"],
                                      "avg_sdf_calls": sts[-1]["sum_sdf"] / sts[-1]["n_pixels"], "max_sdf_calls": sts[-1]["max_sdf"],
                                      "min_sdf_calls": sts[-1]["min_sdf"], "avg_iterations": sts[-1]["sum_iters"] / sts[-1]["n_pixels"]}
                                     for (p, name), sts in zip(SWEEP_PRESETS, last)]}
    # e2e: the same frames through the worker call with host buffers, on a bounded sub-sweep
    if not args.no_e2e and world == 1:
        sub = max(1, min(12, frames))
        t_e = 0.0
        for preset, _ in SWEEP_PRESETS:
            jobs = sweep_jobs(preset, W, H, frames)[:: max(1, frames // sub)][:sub]
            eng = pool if pool_mode else worker
            eng.on_message(jobs[0], shader="normal", pinned=True)
            t1 = time.perf_counter()
            for job in jobs:
                eng.on_message(job, shader="normal", pinned=True)
            t_e += time.perf_counter() - t1
        n_e = sub * len(SWEEP_PRESETS)
        out["e2e"] = {"value": W * H * n_e / t_e / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 120 * n_gpus, "d2h_bytes_per_step": W * H * 12,
                      "ms_per_step": t_e * 1e3 / n_e, "sample": f"{sub} evenly spaced frames per preset, planes into page-locked host memory every frame",
                      "path": "RaymarchPool.on_message" if pool_mode else "RaymarchWorker.on_message -> rm_render"}
    # parity: every 60th frame of every preset, every 16th row, against the oracle (untimed; rank 0)
    bad = False
    if rank == 0 and not args.no_parity:
        from oracle import compare as cmp
        from oracle import pyoracle as po
        chk = rb.RaymarchWorker(device=local_rank) if pool_mode or world > 1 else worker
        rows = np.arange(0, H, 16, dtype=np.int32)
        worst, n_cmp, stats_ok = 1.0, 0, True
        for (preset, _), sts in zip(SWEEP_PRESETS, last):
            jobs = sweep_jobs(preset, W, H, frames)
            s = po.OracleScene().load_preset(preset).build_accel("None")
            for k in (range(59, frames, 60) if frames >= 60 else (frames - 1,)):  # short sweeps: the last frame, never a vacuous pass
                job = jobs[k]
                s.set_camera(job["camera"]["pitch"], job["camera"]["yaw"])
                ref = s.render_rows(W, H, rows, "sphere-tracer")
                f = chk.on_message(job, shader="normal", extras=True)
                a = cmp.fast_agreement(cmp.take_rows(f, W, rows), ref)
                worst = min(worst, a["px_agree"])
                n_cmp += 1
                # the reduced per-frame diagnostics of the timed sweep equal a plain reduction of this frame's planes (main.ts:527-548)
                stats_ok = stats_ok and sts[k]["sum_sdf"] == int(f.sdfEval.astype(np.int64).sum()) and sts[k]["max_sdf"] == int(f.sdfEval.max()) \
                    and sts[k]["sum_iters"] == int(f.iters.astype(np.int64).sum()) and sts[k]["min_sdf"] == int(f.sdfEval.min())
        out["parity"] = {"frames_compared": n_cmp, "rows_per_frame": int(len(rows)), "px_agree_min": worst, "reduced_stats_equal_plane_reductions": stats_ok,
                         "bar": ">= 99.9 % of pixels agree on hit mask, RGB within 1/255, depth rel. err <= 1e-4 (fast build vs oracle)",
                         "pass": bool(n_cmp > 0 and worst >= cmp.PIXEL_AGREEMENT and stats_ok)}
        bad = not out["parity"]["pass"]
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if bad:
        raise SystemExit("bench.py cfg5sweep: parity sample failed")


if __name__ == "__main__":
    main()
