"""End-of-frame pass trigger (RM_TAIL_TRIGGER = parked warps that fire a cooperative pass once a warp's tile queue is empty):
kernel ms / tail ms of cfg4 for the full frame and for the 1/8 stripe share.  python tools/tail_trigger_sweep.py (ONE GPU)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["RM_ANATOMY"] = "1"
import cpu_raymarcher_b200 as rb  # noqa: E402

w = rb.RaymarchWorker(0)
W, H = 3840, 2160
sc = w._ensure_scene(1, "BVH", (100000, 0x5EED0001))


def run(stripes):
    rq = rb.Context.make_request(W, H, sc.camera.get_rotation_matrix3(), sc.camera.get_position(), stripes=stripes, shader="iteration-heatmap")
    best = None
    for _ in range(4):
        w.ctx.render(rq)
        st = w.ctx.stats()
        cur = (st["kernel_ms"], st["drain_ms"], st["tail_ms"], st["tc_passes"])
        best = cur if best is None or cur[0] < best[0] else best
    return best


for trig, order in (("64", "1"), ("4", "1"), ("64", "2"), ("8", "2"), ("4", "2"), ("2", "2")):
    os.environ["RM_TAIL_TRIGGER"] = trig
    os.environ["RM_TILE_ORDER"] = order  # 2: cost-ordered queue for the unstriped full frame too
    full = run(None)
    share = [run((8, 8, i)) for i in (0, 3, 6)]
    print(f"tail trigger {trig:>2} order {order}: full frame {full[0]:.2f} ms (tail {full[2]:.2f}, {full[3]} passes); 1/8 share " +
          " ".join(f"{s[0]:.2f} (tail {s[2]:.2f}, {s[3]} passes)" for s in share), flush=True)
