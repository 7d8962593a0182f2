#!/bin/bash
# quick sweep of every BASELINE config (development aid)
for w in cfg1 cfg2 cfg3 cfg5; do timeout 200 python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1; done
timeout 300 python bench.py --workload cfg4 --prims 10000 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1
timeout 600 python bench.py --workload cfg4 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1
