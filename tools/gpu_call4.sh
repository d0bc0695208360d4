#!/bin/bash
mkdir -p gpurun_out
( time timeout 420 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -12 | tee gpurun_out/gpu_tests_full.log
timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | tee gpurun_out/bench_bwd_c2a.json
timeout 100 python tools/bench_bwd.py c2b 5 2>&1 | tail -1 | tee gpurun_out/bench_bwd_c2b.json
timeout 150 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_c2a.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 30 --csv --log-file gpurun_out/bwd_c2a_launches.csv python tools/bench_bwd.py c2a 2 > /dev/null 2>&1
