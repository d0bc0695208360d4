#!/bin/bash
# round-2 call 27: A-operand collector -- microbenchmark (numerics + timing), then the layered kernel with and without it
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call27.log
: > $LOG
timeout 120 ./tools/mma_collector 2>&1 | tee -a $LOG
one() {  # one <label> <workload> [env...]
  lab=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --workload $wl --steps 200 --warmup 20 --no-cpu-baseline --no-also --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lab', '$wl', 'kernel_us', round(d['roofline']['kernel_us'],2))" | tee -a $LOG
}
one "[plain]" c2b IAF_LY_COLLECTOR=0
one "[collector]" c2b IAF_LY_COLLECTOR=1
one "[plain]" c2b IAF_LY_COLLECTOR=0
one "[collector]" c2b IAF_LY_COLLECTOR=1
one "[plain]" c3 IAF_LY_COLLECTOR=0
one "[collector]" c3 IAF_LY_COLLECTOR=1
IAF_LY_COLLECTOR=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -m gpu -x -q 2>&1 | tail -3 | sed 's/^/[collector=1] /' | tee -a $LOG
exit 0
