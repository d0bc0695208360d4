#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/gpu_tests_full.log
IAF_TC_PREFETCH=4 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_elbo.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/gpu_tests_prefetch4.log
for m in 0 4 0 4 5; do
  IAF_TC_PREFETCH=$m timeout 100 python bench.py --workload c2a --steps 300 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('PREFETCH', $m, round(d['roofline']['kernel_us'],2))" | tee -a gpurun_out/prefetch_jit_ab.log
done
