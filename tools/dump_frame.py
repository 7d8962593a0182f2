"""Render one worker Job through the ctypes path (RaymarchWorker.on_message) and dump the Result planes:
u8 depth | u8 normal | u16 sdfEval | u16 iters — the byte layout tools/node_harness.mjs and tools/ref_fixtures.mjs use.

    python tools/dump_frame.py --dump out.bin --job '{"width":640,"height":360,...}' [--validate]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpu_raymarcher_b200 as rb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dump", required=True)
ap.add_argument("--job", required=True)
ap.add_argument("--validate", action="store_true", help="the fp64 validation build")
a = ap.parse_args()
job = json.loads(a.job)
w = rb.RaymarchWorker(device=0, validate_fp64=a.validate)
f = w.on_message(job)
with open(a.dump, "wb") as fh:
    for plane in (f.depth, f.normal, f.sdfEval, f.iters):
        fh.write(plane.tobytes())
w.close()
