#!/bin/bash
timeout 300 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1
timeout 300 python bench.py --workload cfg4 --prims 10000 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1
timeout 200 python bench.py --workload cfg5 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1
timeout 200 python bench.py --workload cfg3 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1
timeout 200 python bench.py --workload cfg2 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1
