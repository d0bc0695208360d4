#!/bin/bash
# round-2 call 38: recompute on the tensor-core forward; training pair in the bench line; the whole GPU suite
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call38.log
: > $LOG
timeout 400 python -m pytest tests/test_gpu_backward.py -m gpu -x -q 2>&1 | tail -5 | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | cut -c1-600 | sed 's/^/[tc backward] /' | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2b 10 2>&1 | tail -1 | cut -c1-600 | sed 's/^/[tc backward] /' | tee -a $LOG
timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2_bench_default_final2.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_default_final2.json'))
print('default bench: value %.3e  kernel_us %.2f  frac %.3f  e2e %.3e (%.3f ms)  cpu %.3e' % (d['value'], d['roofline']['kernel_us'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline']['value']))
print('also:', json.dumps({k: ({kk: vv for kk, vv in v.items() if kk in ('value','ms_per_step')} if k != 'training_pair' else v) for k, v in d['also'].items()}))" 2>&1 | tee -a $LOG
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a $LOG
exit 0
