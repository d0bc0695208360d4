#!/bin/bash
set -u
mkdir -p gpurun_out
(cd iaf_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DIAF_TC_TIMELINE -shared -Xcompiler -fPIC \
     -o ../lib/libiaf_tl.so iaf_capi.cu iaf_pack.cu iaf_simt.cu iaf_tc.cu iaf_bwd.cu 2>&1 | grep -E "error")
for dbg in 0 15 31; do IAF_FZ_DBG=$dbg timeout 120 python tools/tl_run.py c2a > gpurun_out/r2_tl_fz3_dbg$dbg.log 2>&1; tail -1 gpurun_out/r2_tl_fz3_dbg$dbg.log; done
rm -f iaf_b200/lib/libiaf_tl.so
exit 0
