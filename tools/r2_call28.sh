#!/bin/bash
# round-2 call 28: merged heads in the layered kernel (A/B), the whole GPU suite, the default bench line
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call28.log
: > $LOG
one() {  # one <label> <workload> [env...]
  lab=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --workload $wl --steps 200 --warmup 20 --no-cpu-baseline --no-also --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lab', '$wl', 'kernel_us', round(d['roofline']['kernel_us'],2), 'value', '%.3e' % d['value'])" | tee -a $LOG
}
one "[separate]" c2b IAF_LY_MERGED=0
one "[merged heads]" c2b IAF_LY_MERGED=1
one "[separate]" c2b IAF_LY_MERGED=0
one "[merged heads]" c2b IAF_LY_MERGED=1
one "[separate]" c3 IAF_LY_MERGED=0
one "[merged heads]" c3 IAF_LY_MERGED=1
one "[c2a]" c2a X=1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a $LOG
timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2_bench_default_final.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_default_final.json'))
print('default bench: value %.3e  kernel_us %.2f  frac %.3f  e2e %.3e (%.3f ms)  cpu %.3e  c2b %s' % (d['value'], d['roofline']['kernel_us'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'], json.dumps({k: v for k, v in d['also']['c2b'].items() if k in ('value','ms_per_step')})))" | tee -a $LOG
exit 0
