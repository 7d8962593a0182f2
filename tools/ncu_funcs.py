"""Aggregate an ncu source page per device function of rm_device.cuh (kernel body split in 50-line slices).
Usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_funcs.py"""
import csv
import os
import re
import sys

rows = list(csv.reader(sys.stdin))
hdr = None
cur = None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 9:
        continue
    if r[0] != "":
        cur = int(r[0])
        agg.setdefault(cur, [0.0, 0.0, 0.0])
        continue
    if cur is None:
        continue
    try:
        agg[cur][0] += float(r[hdr.index("Instructions Executed")] or 0)
        agg[cur][1] += float(r[hdr.index("Thread Instructions Executed")] or 0)
        agg[cur][2] += float(r[hdr.index("# Samples")] or 0)
    except ValueError:
        pass
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, "cpu_raymarcher_b200", "csrc", "rm_device.cuh")).read().split("\n")
funcs = []
for i, l in enumerate(src, 1):
    if (l.startswith("RM_DEV") or l.startswith("static __device__") or l.startswith("__global__") or l.startswith("    render_kernel")) and "(" in l:
        name = re.findall(r"(\w+)\(", l)
        if name:
            funcs.append((i, name[0] if name[0] != "__launch_bounds__" else "render_kernel"))


def fn(line):
    prev = "?"
    for i, n in funcs:
        if i > line:
            return prev
        prev = n
    return prev


tot = sum(v[2] for v in agg.values()) or 1
ti = sum(v[0] for v in agg.values()) or 1
byf = {}
for line, v in agg.items():
    f = fn(line)
    if f in ("render_kernel", "?"):
        f = "kernel:%d" % ((line // 50) * 50)
    a = byf.setdefault(f, [0, 0, 0])
    for k in range(3):
        a[k] += v[k]
for f, v in sorted(byf.items(), key=lambda kv: -kv[1][2])[:30]:
    print(f"{f:28s} samp {100*v[2]/tot:5.1f}%  instr {100*v[0]/ti:5.1f}%  lanes {v[1]/max(v[0],1):5.1f}")
