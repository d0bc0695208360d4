#!/bin/bash
# development aid: compare cluster sizes of the layered kernel's weight multicast
for cs in 1 2 4; do
  IAF_LY_CLUSTER=$cs timeout 150 python bench.py --workload ${1:-c2b} --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cs', $cs, d['roofline']['kernel_us'], d['value'], d['roofline']['frac'])"
done
