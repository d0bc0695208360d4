#!/bin/bash
# round-2 call 45: resident head weights in the layered kernel (A/B), parity
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call45.log
: > $LOG
one() {  # one <label> <workload> [env...]
  lab=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --workload $wl --steps 200 --warmup 20 --no-cpu-baseline --no-also --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lab', '$wl', 'kernel_us', round(d['roofline']['kernel_us'],2), 'value', '%.3e' % d['value'])" | tee -a $LOG
}
one "[streamed]" c2b IAF_LY_WRES=0
one "[resident heads]" c2b IAF_LY_WRES=1
one "[streamed]" c2b IAF_LY_WRES=0
one "[resident heads]" c2b IAF_LY_WRES=1
one "[streamed]" c3 IAF_LY_WRES=0
one "[resident heads]" c3 IAF_LY_WRES=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -m gpu -x -q 2>&1 | tail -3 | tee -a $LOG
exit 0
