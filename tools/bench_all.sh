#!/bin/bash
# bench every workload once (N=1), one JSON line each into gpurun_out/bench_all.jsonl
rm -f gpurun_out/bench_all.jsonl
for w in c2a c2b c1 c1_l1 c1_l2 c3 c4_l1; do
  timeout 200 python bench.py --workload $w --steps 200 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 >> gpurun_out/bench_all.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/bench_all.jsonl'):
    try:
        d=json.loads(l)
        print(d['config']['workload'][:60].ljust(62), d['config']['path'], round(d['roofline']['kernel_us'],2), 'us', '%.3e'%d['value'], 'frac', round(d['roofline']['frac'],4))
    except Exception as e:
        print('ERR', l[:200])
PY
