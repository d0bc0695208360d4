#!/bin/bash
# round-2 call 30: context staged by cp.async in the layered kernel's first stage (A/B), the GPU suite
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call30.log
: > $LOG
one() {  # one <label> <workload> [env...]
  lab=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --workload $wl --steps 200 --warmup 20 --no-cpu-baseline --no-also --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lab', '$wl', 'kernel_us', round(d['roofline']['kernel_us'],2), 'value', '%.3e' % d['value'])" | tee -a $LOG
}
one "[ctx by LDG]" c2b IAF_LY_CTXSTAGE=0
one "[ctx by cp.async]" c2b IAF_LY_CTXSTAGE=1
one "[ctx by LDG]" c2b IAF_LY_CTXSTAGE=0
one "[ctx by cp.async]" c2b IAF_LY_CTXSTAGE=1
one "[ctx by LDG]" c3 IAF_LY_CTXSTAGE=0
one "[ctx by cp.async]" c3 IAF_LY_CTXSTAGE=1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a $LOG
exit 0
