#!/bin/bash
# round-2 call 36: TC backward after the lane-parallel copies / shuffle bias sums; the whole GPU suite; sanitizer on the new kernels
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call36.log
: > $LOG
timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -x -q 2>&1 | tail -5 | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | cut -c1-600 | sed 's/^/[tc backward] /' | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2b 10 2>&1 | tail -1 | cut -c1-600 | sed 's/^/[tc backward] /' | tee -a $LOG
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a $LOG
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_backward.py -m gpu -x -q -k "tf-64-16x16 or c2b" 2>&1 | tail -4 | sed 's/^/[memcheck bwd] /' | tee -a $LOG
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_backward.py -m gpu -x -q -k "tf-64-16x16" 2>&1 | tail -4 | sed 's/^/[racecheck bwd] /' | tee -a $LOG
exit 0
