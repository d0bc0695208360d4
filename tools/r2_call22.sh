#!/bin/bash
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call22.log
: > $LOG
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "samples with|passed|failed|Error|FAILED" | cut -c1-500 | sed 's/^/[gpu tests, one box per sample, split weight barrier] /' | tee -a $LOG
one() {  # one <label> <batch> [env...]
  lab=$1; b=$2; shift 2
  env "$@" timeout 200 python bench.py --workload c2a --batch $b --steps 300 --warmup 20 --no-cpu-baseline --no-also --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lab', 'B=$b', 'kernel_us', round(d['roofline']['kernel_us'],2))" | tee -a $LOG
}
one "[staged v3]" 256 X=1; one "[staged v3]" 256 X=1; one "[staged v3]" 32 X=1; one "[staged v3]" 128 X=1
one "[staged v3]" 64 X=1
one "[gathered]" 256 IAF_FZ_STAGE=0
(cd iaf_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DIAF_FZ_PROBE -shared -Xcompiler -fPIC \
     -o ../lib/libiaf_probe.so iaf_capi.cu iaf_pack.cu iaf_simt.cu iaf_tc.cu iaf_bwd.cu 2>&1 | grep -E "error")
echo "== c2a staged x2"; timeout 120 python tools/probe_run.py c2a 2>&1 | grep PROBE | tee gpurun_out/r2_probe6.log
rm -f iaf_b200/lib/libiaf_probe.so
timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | cut -c1-330 | tee -a $LOG
exit 0
