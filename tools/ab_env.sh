#!/bin/bash
# development aid: A/B an environment switch (usage: ab_env.sh VAR workload)
for rep in 1 2; do
  for m in 0 1; do
    env $1=$m timeout 150 python bench.py --workload ${2:-c2a} --steps 300 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', $m, round(d['roofline']['kernel_us'],2))"
  done
done
