#!/bin/bash
# round-2 call 26: A/B of the reduction tail (acq_rel atomic vs fence/atomic/fence) at B = 256 / 64 / 32, then the whole GPU suite
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call26.log
: > $LOG
one() {  # one <label> <batch>
  timeout 200 python bench.py --workload c2a --batch $2 --steps 300 --warmup 20 --no-cpu-baseline --no-also --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'B=$2', 'kernel_us', round(d['roofline']['kernel_us'],2))" | tee -a $LOG
}
for b in 256 64 32; do one "[acq_rel tail]" $b; done
cp iaf_b200/lib/libiaf_b200.so /tmp/lib_keep.so
(cd iaf_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DIAF_FZ_FENCE_TAIL -shared -Xcompiler -fPIC \
     -o ../lib/libiaf_b200.so iaf_capi.cu iaf_pack.cu iaf_simt.cu iaf_tc.cu iaf_bwd.cu 2>&1 | grep -E "error")
for b in 256 64 32; do one "[fence tail]" $b; done
cp /tmp/lib_keep.so iaf_b200/lib/libiaf_b200.so
for b in 256 32; do one "[acq_rel tail again]" $b; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a $LOG
exit 0
