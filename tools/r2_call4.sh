#!/bin/bash
# Round-2 GPU call 4: what slows the MMAs inside the kernel?  (a) microbenchmark: the real stage-1 issue pattern alone and
# against background traffic of each kind; (b) the fz kernel with parts switched off (IAF_FZ_DBG, timing only).
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call4.log
: > $LOG
timeout 200 ./tools/mma_bench 2>&1 | tee gpurun_out/r2_mma_bench3.log | grep time
one() {  # one <label> <workload> [env...]
  lab=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --workload $wl --steps 300 --warmup 20 --no-cpu-baseline --no-also --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lab', '$wl', 'kernel_us', round(d['roofline']['kernel_us'],2))" | tee -a $LOG
}
for dbg in 0 1 3 4 8 12 15 16 19 28 31; do one "[IAF_FZ_DBG=$dbg]" c2a IAF_FZ_DBG=$dbg; done
(cd iaf_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DIAF_TC_TIMELINE -shared -Xcompiler -fPIC \
     -o ../lib/libiaf_tl.so iaf_capi.cu iaf_pack.cu iaf_simt.cu iaf_tc.cu iaf_bwd.cu 2>&1 | grep -E "error")
for dbg in 3 15; do IAF_FZ_DBG=$dbg timeout 120 python tools/tl_run.py c2a > gpurun_out/r2_tl_fz_dbg$dbg.log 2>&1; done
rm -f iaf_b200/lib/libiaf_tl.so
exit 0
