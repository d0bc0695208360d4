#!/bin/bash
# Full validation pass on the GPU box (run under gpurun): the whole -m gpu suite, the training-pair timings, both headline
# bench lines, the launch list of the backward, and smoke().  Outputs land in gpurun_out/.
mkdir -p gpurun_out
( time timeout 420 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -12 | tee gpurun_out/gpu_tests_full.log
timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | tee gpurun_out/bench_bwd_c2a.json
timeout 100 python tools/bench_bwd.py c2b 5 2>&1 | tail -1 | tee gpurun_out/bench_bwd_c2b.json
timeout 150 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_c2a.json
timeout 150 python bench.py --workload c2b --steps 100 --warmup 10 2>&1 | tail -1 | tee gpurun_out/bench_c2b.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/bwd_c2a_launches.csv python tools/bench_bwd.py c2a 2 > /dev/null 2>&1
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -8 | tee gpurun_out/smoke.log
