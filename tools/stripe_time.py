"""Kernel time of config 4 when one GPU renders only the stripes of rank i of N (what each GPU of an N-GPU run does), for every
i: isolates the end-of-frame tail (time above ideal = full frame / N) and the skew between ranks (max - mean) from the multi-GPU
plumbing.  Stripe heights 4 and 8 rows, cost-ordered tile queue on and off.  python tools/stripe_time.py (needs ONE GPU)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["RM_ANATOMY"] = "1"
import cpu_raymarcher_b200 as rb  # noqa: E402

w = rb.RaymarchWorker(0)
W, H = 3840, 2160
sc = w._ensure_scene(1, "BVH", (100000, 0x5EED0001))


def kernel_ms(stripes):
    rq = rb.Context.make_request(W, H, sc.camera.get_rotation_matrix3(), sc.camera.get_position(), stripes=stripes, shader="iteration-heatmap")
    ts = []
    for _ in range(4):
        w.ctx.render(rq)
        st = w.ctx.stats()
        ts.append((st["kernel_ms"], st["drain_ms"], st["tail_ms"]))
    best = min(ts[1:])
    kernel_ms.anatomy.append(best)
    return best[0]


kernel_ms.anatomy = []


for order in ("1", "0"):
    os.environ["RM_TILE_ORDER"] = order
    full = kernel_ms(None)
    a = kernel_ms.anatomy[-1]
    print(f"tile order {'on' if order == '1' else 'off'}: full frame {full:.2f} ms (queue drained after {a[1]:.2f} ms, tail {a[2]:.2f} ms)")
    for rows in (4, 8):
        for n in ((2, 4, 8) if rows == 8 else (8,)):
            ts = [kernel_ms((rows, n, i)) for i in range(n)]
            an = kernel_ms.anatomy[-n:]
            print(f"  {rows}-row stripes, 1/{n} of the frame: ideal {full / n:.2f} ms; per rank " + " ".join(f"{t:.2f}" for t in ts) +
                  f"; mean {sum(ts) / n:.2f} max {max(ts):.2f} -> efficiency {full / n / max(ts):.3f}; queue drained after " +
                  f"{sum(a[1] for a in an) / n:.2f} ms, tail {sum(a[2] for a in an) / n:.2f} ms (mean over ranks)")
