#!/bin/bash
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call12.log
: > $LOG
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_elbo.py tests/test_elbo_theano.py -m gpu -x -q 2>&1 | tail -2 | sed 's/^/[gpu parity+elbo] /' | tee -a $LOG
one() {  # one <label> <batch> [env...]
  lab=$1; b=$2; shift 2
  env "$@" timeout 200 python bench.py --workload c2a --batch $b --steps 300 --warmup 20 --no-cpu-baseline --no-also --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lab', 'B=$b', 'kernel_us', round(d['roofline']['kernel_us'],2), 'clk', d['clocks']['sm_mhz'])" | tee -a $LOG
}
one "[real]" 256 X=1; one "[real]" 256 X=1; one "[real]" 32 X=1
for dbg in 31 15 28 12; do one "[IAF_FZ_DBG=$dbg]" 256 IAF_FZ_DBG=$dbg; done
one "[IAF_FZ_DBG=31]" 32 IAF_FZ_DBG=31
exit 0
