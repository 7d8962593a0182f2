"""Per-launch facts of one kernel from an .ncu-rep, as the JSON bench.py reads for its roofline block (profiles/ncu_*.json).

    python tools/ncu_facts.py gpurun_out/X.ncu-rep <workload> <n_prims> <W> <H> profiles/ncu_<round>_<workload>.json

Runs on the CPU box (ncu -i).  Fields: traffic_bytes (dram read + write), issue_util (smsp__issue_active / peak), lanes_per_inst
(smsp__thread_inst_executed_per_inst_executed), warp instructions, kernel time under ncu, the five largest warp-stall reasons."""
import csv
import io
import json
import subprocess
import sys

rep, wl, n_prims, W, H, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
d = dict(zip(hdr, r))
u = dict(zip(hdr, units))


def num(k):
    return float(d[k].replace(",", ""))


def to_bytes(k):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[k]]
    return num(k) * scale


stalls = {k.split("issue_stalled_")[1].split("_per_issue_active")[0]: num(k) for k in hdr
          if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
tot = sum(stalls.values()) or 1.0
top = sorted(stalls.items(), key=lambda kv: -kv[1])[:5]
ms = num("gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u["gpu__time_duration.sum"]]
facts = {
    "workload": wl, "n_prims": n_prims, "width": W, "height": H, "kernel": d.get("Kernel Name"),
    "source": f"{rep.split('/')[-1]} (ncu --set full --clock-control none, one launch)",
    "kernel_ms_under_ncu": ms,
    "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
    "traffic_bytes": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
    "issue_util": num("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
    "lanes_per_inst": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "warp_instructions": num("smsp__inst_executed.sum"),
    "registers_per_thread": num("launch__registers_per_thread"),
    "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "pipe_fma_pct": num("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    "stalls_top": [{"reason": k, "share": v / tot} for k, v in top],
}
json.dump(facts, open(out, "w"), indent=1)
print(json.dumps(facts))
