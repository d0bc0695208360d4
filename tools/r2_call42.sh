#!/bin/bash
# round-2 call 42: final validation of the committed tree at N=1 -- smoke, the whole GPU suite, both bench arms
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call42.log
: > $LOG
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a $LOG
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a $LOG
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-400 | sed 's/^/[reference arm] /' | tee -a $LOG
timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2_bench_default_final3.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_default_final3.json'))
print('default bench: value %.3e  kernel_us %.2f  frac %.3f  e2e %.3e (%.3f ms)  cpu %.3e  launches %s  clocks %s' % (d['value'], d['roofline']['kernel_us'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'], d.get('gpu_launches'), json.dumps(d.get('clocks'))))
print('also:', json.dumps({k: ({kk: vv for kk, vv in v.items() if kk in ('value','ms_per_step')} if k != 'training_pair' else v) for k, v in d['also'].items()}))" 2>&1 | tee -a $LOG
exit 0
