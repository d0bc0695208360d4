#!/bin/bash
# development aid: timing experiments of the layered kernel (IAF_LY_DBG breaks the results on purpose)
for d in ${DBGS:-0 1 2 4 6 7}; do
  IAF_LY_DBG=$d timeout 150 python bench.py --workload ${1:-c2b} --steps 200 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dbg', $d, round(d['roofline']['kernel_us'],2))"
done
