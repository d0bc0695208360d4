#!/bin/bash
# round-2 call 47: last validation of the committed tree -- smoke, the whole GPU suite, the default bench line
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call47.log
: > $LOG
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 | tee -a $LOG
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a $LOG
timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2_bench_default_final4.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_default_final4.json'))
print('default bench: value %.3e  kernel_us %.2f  frac %.3f  e2e %.3e (%.3f ms)  cpu %.3e (%s thr)' % (d['value'], d['roofline']['kernel_us'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'], d['cpu_baseline']['cores']))
print('also:', json.dumps({k: ({kk: vv for kk, vv in v.items() if kk in ('value','ms_per_step')} if k != 'training_pair' else v) for k, v in d['also'].items()}))" 2>&1 | tee -a $LOG
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('[reference arm] value %.3e  ms/step %.2f  cores %s  candidates %s' % (d['value'], d['ms_per_step'], d['cpu_baseline']['cores'], d['cpu_baseline']['candidates_ms']))" 2>&1 | tee -a $LOG
exit 0
