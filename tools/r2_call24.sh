#!/bin/bash
set -u
mkdir -p gpurun_out
(cd iaf_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DIAF_FZ_PROBE -shared -Xcompiler -fPIC \
     -o ../lib/libiaf_probe.so iaf_capi.cu iaf_pack.cu iaf_simt.cu iaf_tc.cu iaf_bwd.cu 2>&1 | grep -E "error")
for cfg in "X=1" "IAF_FZ_DBG=1" "IAF_FZ_TWO_STAGE=1" "IAF_FZ_TWO_STAGE=1 IAF_FZ_DBG=1" "IAF_FZ_STAGE=0"; do echo "== $cfg"; env $cfg timeout 120 python tools/probe_run.py c2a 2>&1 | grep PROBE; done | tee gpurun_out/r2_probe7.log
rm -f iaf_b200/lib/libiaf_probe.so
one() {  # one <label> <batch> [env...]
  lab=$1; b=$2; shift 2
  env "$@" timeout 200 python bench.py --workload c2a --batch $b --steps 300 --warmup 20 --no-cpu-baseline --no-also --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lab', 'B=$b', 'kernel_us', round(d['roofline']['kernel_us'],2))"
}
one "[two stage v3]" 256 IAF_FZ_TWO_STAGE=1
exit 0
