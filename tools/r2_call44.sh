#!/bin/bash
# round-2 call 44: M = 64 weight-gradient blocks -- backward tests, timing
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call44.log
: > $LOG
timeout 400 python -m pytest tests/test_gpu_backward.py -m gpu -x -q 2>&1 | tail -25 | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | cut -c1-600 | sed 's/^/[M=64 blocks] /' | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2b 10 2>&1 | tail -1 | cut -c1-600 | sed 's/^/[M=64 blocks] /' | tee -a $LOG
exit 0
