#!/bin/bash
# round-2 call 48: bench line with the one-graph timed region at N=1
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call48.log
: > $LOG
for rep in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2_bench_default_final5.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_default_final5.json'))
print('default bench: value %.3e (%.2f us/step)  kernel_us %.2f  frac %.3f  e2e %.3e (%.3f ms) launches %s' % (d['value'], d['ms_per_step']*1e3, d['roofline']['kernel_us'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches']))
print('also:', json.dumps({k: ({kk: vv for kk, vv in v.items() if kk in ('value','ms_per_step')} if k != 'training_pair' else v) for k, v in d['also'].items()}))" 2>&1 | tee -a $LOG
done
timeout 300 python bench.py --steps 400 --warmup 20 --no-cpu-baseline --no-also --no-e2e 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('400 steps: value %.3e (%.2f us/step) kernel_us %.2f' % (d['value'], d['ms_per_step']*1e3, d['roofline']['kernel_us']))" | tee -a $LOG
exit 0
