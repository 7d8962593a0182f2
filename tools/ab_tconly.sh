#!/bin/bash
# Build first: make -C cpu_raymarcher_b200/csrc exp-tconly
# In-call A/B on one box: shipped library vs the RM_EXP_TC_ONLY build (tensor-core kernel without its FFMA fallback body).
run() { timeout 40 python bench.py --no-cpu-baseline --no-e2e --steps 6 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), 'ms')"; }
cp cpu_raymarcher_b200/librm_b200.so /tmp/ship.so
run A_shipped
cp build/exp/librm_b200_tconly.so cpu_raymarcher_b200/librm_b200.so
run B_tconly
run B_tconly
cp /tmp/ship.so cpu_raymarcher_b200/librm_b200.so
run A_shipped
