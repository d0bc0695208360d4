#!/bin/bash
# one GPU call: backward parity tests, prefetch A/B, backward timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_backward.py tests/test_elbo.py -m gpu -q 2>&1 | tail -25 > gpurun_out/bwd_tests.log
cat gpurun_out/bwd_tests.log
for m in 0 1 2 0 1 2; do
  IAF_TC_PREFETCH=$m timeout 100 python bench.py --workload c2a --steps 300 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('PREFETCH', $m, round(d['roofline']['kernel_us'],2))" | tee -a gpurun_out/prefetch_ab.log
done
timeout 120 python tools/bench_bwd.py c2a 20 2>&1 | tail -2 | tee gpurun_out/bench_bwd_c2a.json
timeout 120 python tools/bench_bwd.py c2b 5 2>&1 | tail -2 | tee gpurun_out/bench_bwd_c2b.json
