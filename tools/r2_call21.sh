#!/bin/bash
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call21.log
: > $LOG
bash tools/profile.sh c2a r2 > /dev/null 2>&1
bash tools/profile.sh c2b r2 > /dev/null 2>&1
one() {  # one <label> <workload> [env...]
  lab=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --workload $wl --steps 200 --warmup 20 --no-cpu-baseline --no-also --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lab', '$wl', 'kernel_us', round(d['roofline']['kernel_us'],2), 'frac', round(d['roofline']['frac'],3), d['roofline']['bound'], 'path', d['config']['path'])" | tee -a $LOG
}
for wl in c2a c2b c1 c1_l1 c1_l2 c3 c4_l1; do one "[r2 final]" $wl X=1; done
timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | tee gpurun_out/r2_bench_bwd_c2a.json | cut -c1-400 | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2b 10 2>&1 | tail -1 | tee gpurun_out/r2_bench_bwd_c2b.json | cut -c1-400 | tee -a $LOG
ls -la gpurun_out/*.ncu-rep | tee -a $LOG
exit 0
