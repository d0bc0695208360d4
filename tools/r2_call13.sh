#!/bin/bash
set -u
mkdir -p gpurun_out
(cd iaf_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DIAF_FZ_PROBE -shared -Xcompiler -fPIC \
     -o ../lib/libiaf_probe.so iaf_capi.cu iaf_pack.cu iaf_simt.cu iaf_tc.cu iaf_bwd.cu 2>&1 | grep -E "error")
for dbg in 0 15 31 12 28; do echo "== IAF_FZ_DBG=$dbg"; IAF_FZ_DBG=$dbg timeout 120 python tools/probe_run.py c2a 2>&1 | grep PROBE; done | tee gpurun_out/r2_probe.log
rm -f iaf_b200/lib/libiaf_probe.so
exit 0
