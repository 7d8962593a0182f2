#!/bin/bash
# Multi-GPU measurement set for one box with N GPUs (run under gpurun --gpus N): both routes of DESIGN.md §5 at every
# power-of-two GPU count up to N, the multi-GPU equality checks, and the config-5 sweep.  Output: gpurun_out/<tag>_*.json / .log
#   tools/scale_runs.sh <tag> <N> [sweep_frames]
TAG=${1:-r02} ; N=${2:-2} ; FR=${3:-360}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# 1. equality: pool spanning real devices + the torchrun route (frames and stats == single GPU, bit for bit)
(python -m pytest tests/test_gpu_pool.py "tests/test_gpu_parity.py::test_multi_gpu_fused_gather_matches_single_gpu" -m gpu -x -q 2>&1 | tail -5) > gpurun_out/${TAG}_multigpu_tests_${N}gpu.log 2>&1
$TR --nproc-per-node $N --master-port 29611 tools/multigpu_check.py > gpurun_out/${TAG}_multigpu_check_${N}gpu.log 2>&1
# 2. headline workload at 1, 2, 4, ... N GPUs, both routes
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_1gpu_of${N}.json 2>> gpurun_out/${TAG}_err.log
for g in 2 4 8; do
  [ $g -le $N ] || continue
  $TR --nproc-per-node $g --master-port $((29620 + g)) bench.py --gpus $g --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_${g}gpu_torchrun.json 2>> gpurun_out/${TAG}_err.log
  python bench.py --gpus $g --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_${g}gpu_pool.json 2>> gpurun_out/${TAG}_err.log
done
# 3. config 5: the analytics rotation sweep on all N GPUs (frame-parallel and stripe-parallel, both routes)
python bench.py --workload cfg5sweep --gpus $N --frames $FR --steps 1 --warmup 1 --sweep-mode frames > gpurun_out/${TAG}_sweep_${N}gpu_pool_frames.json 2>> gpurun_out/${TAG}_err.log
python bench.py --workload cfg5sweep --gpus $N --frames $FR --steps 1 --warmup 1 --sweep-mode stripes --no-parity > gpurun_out/${TAG}_sweep_${N}gpu_pool_stripes.json 2>> gpurun_out/${TAG}_err.log
$TR --nproc-per-node $N --master-port 29640 bench.py --workload cfg5sweep --gpus $N --frames $FR --steps 1 --warmup 1 --sweep-mode frames --no-parity > gpurun_out/${TAG}_sweep_${N}gpu_torchrun_frames.json 2>> gpurun_out/${TAG}_err.log
tail -3 gpurun_out/${TAG}_multigpu_tests_${N}gpu.log; tail -12 gpurun_out/${TAG}_multigpu_check_${N}gpu.log
for f in gpurun_out/${TAG}_bench_*gpu*.json gpurun_out/${TAG}_sweep_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:b.get(k) for k in ("n_gpus","value","ms_per_step","wall_ms_per_step","kernel_ms_per_gpu","frames_per_s","scene_upload_ms")}, "e2e", (b.get("e2e") or {}).get("value"), (b.get("e2e") or {}).get("ms_per_step"), "parity", (b.get("parity") or {}).get("pass"))
except Exception as e:
    print("unreadable:", e)
PY
done
tail -5 gpurun_out/${TAG}_err.log
