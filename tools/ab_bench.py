"""In-call A/B of library variants (boxes of the pool differ by up to ~9 %, so only numbers from ONE gpurun call compare).

    python tools/ab_bench.py [--rounds 2] [--steps 5] name=path/to/lib.so ... -- <bench.py args>

Runs bench.py (device-resident value only: --no-e2e --no-cpu-baseline --no-parity) once per variant per round, interleaved,
and prints ms_per_step per variant.  `default` = the shipped cpu_raymarcher_b200/librm_b200.so."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    argv = sys.argv[1:]
    extra = []
    if "--" in argv:
        i = argv.index("--")
        argv, extra = argv[:i], argv[i + 1:]
    rounds, steps, variants = 2, 5, [("default", None)]
    it = iter(argv)
    for a in it:
        if a == "--rounds":
            rounds = int(next(it))
        elif a == "--steps":
            steps = int(next(it))
        else:
            name, path = a.split("=", 1)
            variants.append((name, path))
    res = {n: [] for n, _ in variants}
    for _ in range(rounds):
        for name, path in variants:
            env = dict(os.environ)
            if path:
                env["RM_B200_LIB"] = os.path.join(ROOT, path)
            else:
                env.pop("RM_B200_LIB", None)
            out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(steps), "--warmup", "3", "--no-e2e",
                                  "--no-cpu-baseline", "--no-parity"] + extra, env=env, capture_output=True, text=True)
            try:
                line = json.loads(out.stdout.strip().splitlines()[-1])
                res[name].append(round(line["ms_per_step"], 3))
            except Exception:
                res[name].append("FAILED: " + (out.stderr.strip().splitlines() or ["?"])[-1][:200])
            print(json.dumps({"variant": name, "ms_per_step": res[name][-1]}), flush=True)  # per run: a call cut off by its time limit still reports
    print(json.dumps({"args": extra, "ms_per_step": res}), flush=True)


if __name__ == "__main__":
    main()
