#!/bin/bash
# development aid: build the library with different worker-warp counts on the GPU box and bench each
cd iaf_b200/csrc
for w in 16 8; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DTC_WORKERS=$w -shared -Xcompiler -fPIC \
    -o ../lib/libiaf_b200.so iaf_capi.cu iaf_pack.cu iaf_simt.cu iaf_tc.cu iaf_bwd.cu 2>&1 | grep -E "error"
  for rep in 1 2; do
    (cd ../.. && timeout 150 python bench.py --workload ${1:-c2a} --steps 300 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('workers', $w, round(d['roofline']['kernel_us'],2))")
  done
done
(cd ../.. && timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2)
