#!/bin/bash
# Round-2 GPU call 7: fz kernel with decoupled loaders (two z windows, context first, tables by bulk copy)
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call7.log
: > $LOG
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | sed 's/^/[gpu tests] /' | tee -a $LOG
one() {  # one <label> <batch> [env...]
  lab=$1; b=$2; shift 2
  env "$@" timeout 200 python bench.py --workload c2a --batch $b --steps 300 --warmup 20 --no-cpu-baseline --no-also --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lab', 'B=$b', 'kernel_us', round(d['roofline']['kernel_us'],2))" | tee -a $LOG
}
one "[real]" 256 X=1; one "[real]" 256 X=1; one "[real]" 32 X=1
for dbg in 31 15 28 16 12 8 4 3; do one "[IAF_FZ_DBG=$dbg]" 256 IAF_FZ_DBG=$dbg; done
one "[IAF_FZ_DBG=31]" 32 IAF_FZ_DBG=31
(cd iaf_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DIAF_TC_TIMELINE -shared -Xcompiler -fPIC \
     -o ../lib/libiaf_tl.so iaf_capi.cu iaf_pack.cu iaf_simt.cu iaf_tc.cu iaf_bwd.cu 2>&1 | grep -E "error")
timeout 120 python tools/tl_run.py c2a > gpurun_out/r2_tl_fz2.log 2>&1
rm -f iaf_b200/lib/libiaf_tl.so
exit 0
