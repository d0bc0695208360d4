#!/bin/bash
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call14.log
: > $LOG
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | sed 's/^/[gpu tests] /' | tee -a $LOG
one() {  # one <label> <batch> [env...]
  lab=$1; b=$2; shift 2
  env "$@" timeout 200 python bench.py --workload c2a --batch $b --steps 300 --warmup 20 --no-cpu-baseline --no-also --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lab', 'B=$b', 'kernel_us', round(d['roofline']['kernel_us'],2))" | tee -a $LOG
}
one "[staged]" 256 X=1; one "[staged]" 256 X=1; one "[staged]" 32 X=1
one "[gathered]" 256 IAF_FZ_STAGE=0; one "[gathered]" 32 IAF_FZ_STAGE=0
for dbg in 31 15 12; do one "[staged IAF_FZ_DBG=$dbg]" 256 IAF_FZ_DBG=$dbg; done
(cd iaf_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DIAF_FZ_PROBE -shared -Xcompiler -fPIC \
     -o ../lib/libiaf_probe.so iaf_capi.cu iaf_pack.cu iaf_simt.cu iaf_tc.cu iaf_bwd.cu 2>&1 | grep -E "error")
for dbg in 0; do echo "== IAF_FZ_DBG=$dbg"; IAF_FZ_DBG=$dbg timeout 120 python tools/probe_run.py c2a 2>&1 | grep PROBE; done | tee gpurun_out/r2_probe2.log
rm -f iaf_b200/lib/libiaf_probe.so
exit 0
