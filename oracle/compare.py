"""ORACLE — TEST INFRASTRUCTURE ONLY.  Parity metrics between a frame of the CUDA path and the oracle's.

Shared by tests/ and by bench.py's `parity` block (the checker leg; never the thing measured).  Both arguments
are frame-like objects with the reference's Result planes (raymarchWorker.ts:24-31) — `depth` u8, `normal` u8x3,
`sdfEval` u16, `iters` u16 — plus `depth_f64` (the unquantised rayMarch return); the oracle's frame also
carries `sdf_full` / `iters_full`, the GPU's `sdf_u32`.

Bars (BASELINE.json north_star):
  * fp64 validation build: hit mask, SDF-call and iteration counters bit-exact (here also: depth/normal bytes and the
    unquantised depth as raw doubles);
  * fp32 fast path: >= 99.9 % of pixels agree on the hit mask AND are within 1/255 in RGB AND have a depth
    relative error <= 1e-4.
"""
from __future__ import annotations

import numpy as np

RGB_TOL = 1            # 1/255
DEPTH_REL_TOL = 1e-4
PIXEL_AGREEMENT = 0.999
MAX_DIST = 10.0


def take_rows(frame, width: int, rows, y_start: int = 0):
    """The planes of image rows `rows` of a band frame that starts at image row y_start (compact, like OracleScene.render_rows)."""
    rows = np.asarray(rows, np.int64) - y_start
    idx = (rows[:, None] * width + np.arange(width)[None, :]).reshape(-1)

    class Rows:
        pass

    out = Rows()
    for name in ("depth", "sdfEval", "iters", "depth_f32", "sdf_u32", "depth_f64", "sdf_full", "iters_full"):
        v = getattr(frame, name, None)
        setattr(out, name, None if v is None else np.asarray(v)[idx])
    out.normal = np.asarray(frame.normal).reshape(-1, 3)[idx].reshape(-1)
    for name in ("rgba", "rgba_analytics"):
        v = getattr(frame, name, None)
        setattr(out, name, None if v is None else np.asarray(v).reshape(-1, 4)[idx].reshape(-1))
    return out


def bit_exact(gpu, ref) -> dict:
    """Validation-build bar: every plane identical.  Returns per-plane booleans and `all`."""
    out = {
        "sdf_counters": bool(np.array_equal(gpu.sdfEval, ref.sdfEval)),
        "iter_counters": bool(np.array_equal(gpu.iters, ref.iters)),
        "hit_mask": bool(np.array_equal(gpu.depth_f64 < MAX_DIST, ref.depth_f64 < MAX_DIST)),
        "depth_bits": bool(np.array_equal(np.asarray(gpu.depth_f64).view(np.uint64), np.asarray(ref.depth_f64).view(np.uint64))),
        "depth_bytes": bool(np.array_equal(gpu.depth, ref.depth)),
        "normal_bytes": bool(np.array_equal(gpu.normal, ref.normal)),
    }
    if getattr(gpu, "sdf_u32", None) is not None and getattr(ref, "sdf_full", None) is not None:
        out["sdf_counters_unwrapped"] = bool(np.array_equal(gpu.sdf_u32, ref.sdf_full))
    out["all"] = all(out.values())
    return out


def fast_agreement(gpu, ref, rgba_ref=None) -> dict:
    """Fast-path bar, literally.  `rgba_ref` (optional) = the oracle's shade of its own planes with the shader the GPU frame's
    `rgba` was produced with; it joins the RGB test."""
    hit_ref, hit_gpu = ref.depth_f64 < MAX_DIST, gpu.depth_f64 < MAX_DIST
    hit_ok = hit_ref == hit_gpu
    dn = np.abs(np.asarray(gpu.normal).reshape(-1, 3).astype(np.int16) - np.asarray(ref.normal).reshape(-1, 3).astype(np.int16)).max(1)
    rgb = dn.copy()
    if rgba_ref is not None and getattr(gpu, "rgba", None) is not None:
        dc = np.abs(np.asarray(gpu.rgba).reshape(-1, 4).astype(np.int16) - np.asarray(rgba_ref).reshape(-1, 4).astype(np.int16)).max(1)
        rgb = np.maximum(rgb, dc)
    rel = np.abs(gpu.depth_f64 - ref.depth_f64) / np.maximum(np.abs(ref.depth_f64), 1e-12)
    ok = hit_ok & (rgb <= RGB_TOL) & (rel <= DEPTH_REL_TOL)
    cnt = (np.asarray(gpu.sdfEval) == np.asarray(ref.sdfEval)) & (np.asarray(gpu.iters) == np.asarray(ref.iters))
    n = int(ok.size)
    return {
        "pixels": n,
        "px_agree": float(ok.mean()) if n else 1.0,
        "hit_agree": float(hit_ok.mean()) if n else 1.0,
        "rgb_within_1": float((rgb <= RGB_TOL).mean()) if n else 1.0,
        "rgb_max": int(rgb.max()) if n else 0,
        "depth_within_1e-4": float((rel <= DEPTH_REL_TOL).mean()) if n else 1.0,
        "depth_rel_max": float(rel.max()) if n else 0.0,
        "counters_equal": float(cnt.mean()) if n else 1.0,
        "pass": bool((float(ok.mean()) if n else 1.0) >= PIXEL_AGREEMENT),
    }
