#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_backward.py tests/test_elbo.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/bwd_tests2.log
timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | tee gpurun_out/bench_bwd_c2a.json
timeout 100 python tools/bench_bwd.py c2b 5 2>&1 | tail -1 | tee gpurun_out/bench_bwd_c2b.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'iaf_lconv|iaf_bwd_wgrad' -s 30 -c 5 -o gpurun_out/bwd_c2a -f python tools/bench_bwd.py c2a 1 > gpurun_out/ncu_bwd.log 2>&1
tail -3 gpurun_out/ncu_bwd.log
