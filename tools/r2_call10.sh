#!/bin/bash
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call10.log
: > $LOG
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2 | sed 's/^/[gpu parity] /' | tee -a $LOG
one() {  # one <label> <batch> [env...]
  lab=$1; b=$2; shift 2
  env "$@" timeout 200 python bench.py --workload c2a --batch $b --steps 300 --warmup 20 --no-cpu-baseline --no-also --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lab', 'B=$b', 'kernel_us', round(d['roofline']['kernel_us'],2))" | tee -a $LOG
}
one "[real]" 256 X=1; one "[real]" 256 X=1; one "[real]" 32 X=1
for dbg in 15 31; do one "[IAF_FZ_DBG=$dbg]" 256 IAF_FZ_DBG=$dbg; done
(cd iaf_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DIAF_TC_TIMELINE -shared -Xcompiler -fPIC \
     -o ../lib/libiaf_tl.so iaf_capi.cu iaf_pack.cu iaf_simt.cu iaf_tc.cu iaf_bwd.cu 2>&1 | grep -E "error")
for dbg in 0 15; do IAF_FZ_DBG=$dbg timeout 120 python tools/tl_run.py c2a > gpurun_out/r2_tl_fz4_dbg$dbg.log 2>&1; done
rm -f iaf_b200/lib/libiaf_tl.so
exit 0
