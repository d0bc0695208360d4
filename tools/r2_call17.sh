#!/bin/bash
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call17.log
: > $LOG
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "samples with|passed|failed|Error|FAILED" | cut -c1-500 | sed 's/^/[gpu tests fp16 operands] /' | tee -a $LOG
one() {  # one <label> <workload> [env...]
  lab=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --workload $wl --steps 200 --warmup 20 --no-cpu-baseline --no-also --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lab', '$wl', 'kernel_us', round(d['roofline']['kernel_us'],2), 'frac', round(d['roofline']['frac'],3))" | tee -a $LOG
}
one "[gathered]" c2a IAF_FZ_STAGE=0
one "[c2b]" c2b X=1
one "[c2b cluster2]" c2b IAF_LY_CLUSTER=2
one "[c3 B=32]" c3 X=1
one "[c3 B=32 cluster2]" c3 IAF_LY_CLUSTER=2
# the driver's own invocation, default flags
timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2_bench_default.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_default.json'))
print('BENCH value %.3e ms/step %.4f kernel_us %.2f frac %.3f bound %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_us'], d['roofline']['frac'], d['roofline']['bound']))
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['host_binding'])
print('cpu', d['cpu_baseline'])
print('also', json.dumps(d['also'])[:600])
print('clocks', d['clocks'])
" | tee -a $LOG
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2_bench_reference.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_reference.json'))
print('REFERENCE value %.3e ms/step %.3f cores %s cands %s' % (d['value'], d['ms_per_step'], d['cpu_baseline']['cores'], d['cpu_baseline']['candidates_ms']))
" | tee -a $LOG
exit 0
