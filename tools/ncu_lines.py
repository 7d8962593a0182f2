"""Summarise an ncu source page per CUDA source line.
Usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_lines.py [N]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 25
hdr = None
cur = None
agg = {}
order = []
for r in rows:
    if not r:
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 9:
        continue
    if r[0] != "":
        cur = (r[0], r[1].strip()[:100])
        if cur not in agg:
            agg[cur] = [0.0, 0.0, 0.0]
            order.append(cur)
        continue
    if cur is None:
        continue
    i_inst = hdr.index("Instructions Executed")
    i_thr = hdr.index("Thread Instructions Executed")
    i_s = hdr.index("# Samples")
    try:
        agg[cur][0] += float(r[i_inst] or 0)
        agg[cur][1] += float(r[i_thr] or 0)
        agg[cur][2] += float(r[i_s] or 0)
    except ValueError:
        pass
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[2] for v in agg.values()) or 1
print(f"total warp-instr {ti:.3e}  lanes/instr {sum(v[1] for v in agg.values())/ti:.2f}  samples {ts:.0f}")
print("--- top by stall samples (~time)")
for k in sorted(agg, key=lambda k: -agg[k][2])[:N]:
    v = agg[k]
    print(f"L{k[0]:>4s} samp {100*v[2]/ts:5.1f}%  instr {100*v[0]/ti:5.1f}%  lanes {v[1]/max(v[0],1):5.1f}  {k[1]}")
