"""rm_render (host planes, page-locked: early band download) on the 1/8 stripe share of cfg4: wall time of the call vs kernel time,
with the cost-ordered tile queue on and off.  What one GPU of an 8-GPU end-to-end frame does.  python tools/share_e2e.py (ONE GPU)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["RM_ANATOMY"] = "1"
import cpu_raymarcher_b200 as rb  # noqa: E402

w = rb.RaymarchWorker(0)
W, H = 3840, 2160
sc = w._ensure_scene(1, "BVH", (100000, 0x5EED0001))
for order in ("1", "0"):
    os.environ["RM_TILE_ORDER"] = order
    for n in (8, 4, 1):
        res = []
        for i in (range(n) if n > 1 else (0,)):
            rq = rb.Context.make_request(W, H, sc.camera.get_rotation_matrix3(), sc.camera.get_position(), stripes=(8, n, i) if n > 1 else None,
                                         shader="iteration-heatmap")
            best = None
            for _ in range(5):
                t0 = time.perf_counter()
                w.ctx.render(rq, pinned=True)
                wall = (time.perf_counter() - t0) * 1e3
                st = w.ctx.stats()
                cur = (wall, st["kernel_ms"], st["tail_ms"])
                best = cur if best is None or cur[0] < best[0] else best
            res.append(best)
        print(f"tile order {order}, 1/{n}: rm_render wall ms " + " ".join(f"{r[0]:.2f}" for r in res) + " | kernel ms " + " ".join(f"{r[1]:.2f}" for r in res) +
              f" | mean wall - kernel = {sum(r[0] - r[1] for r in res) / len(res):.2f} ms", flush=True)
