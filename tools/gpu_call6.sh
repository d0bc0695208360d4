#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/gpu_tests_full.log
timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | tee gpurun_out/bench_bwd_c2a.json
timeout 100 python tools/bench_bwd.py c2b 5 2>&1 | tail -1 | tee gpurun_out/bench_bwd_c2b.json
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 24 --csv --log-file gpurun_out/bwd_c2a_launches.csv python tools/bench_bwd.py c2a 2 > /dev/null 2>&1
