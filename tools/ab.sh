#!/bin/bash
# development aid: A/B two env settings of the fused kernel on the same box
for rep in 1 2; do
  for m in 0 1; do
    IAF_TC_MERGED=$m timeout 150 python bench.py --workload ${1:-c2a} --steps 300 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('merged', $m, round(d['roofline']['kernel_us'],2))"
  done
done
