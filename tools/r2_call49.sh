#!/bin/bash
# round-2 call 49: weight-gradient kernel builds its X tiles itself (no input image) -- backward tests, timing, whole suite
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call49.log
: > $LOG
timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -x -q 2>&1 | tail -15 | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | cut -c1-700 | sed 's/^/[x in kernel] /' | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2b 10 2>&1 | tail -1 | cut -c1-700 | sed 's/^/[x in kernel] /' | tee -a $LOG
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a $LOG
timeout 200 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_backward.py -m gpu -x -q -k "tf-64-16x16" 2>&1 | tail -3 | sed 's/^/[racecheck] /' | tee -a $LOG
exit 0
