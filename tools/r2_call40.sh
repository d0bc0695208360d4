#!/bin/bash
# round-2 call 40: backward tests (batch-shrink case), sanitizers on the final tensor-core backward, ncu of its two MMA kernels
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call40.log
: > $LOG
timeout 400 python -m pytest tests/test_gpu_backward.py -m gpu -x -q 2>&1 | tail -5 | tee -a $LOG
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_backward.py -m gpu -x -q -k "tensor_core or tf-64-16x16 or theano-160x160" 2>&1 | tail -4 | sed 's/^/[memcheck tc backward] /' | tee -a $LOG
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_backward.py -m gpu -x -q -k "tf-64-16x16 or theano-64-8x8" 2>&1 | tail -4 | sed 's/^/[racecheck tc backward] /' | tee -a $LOG
timeout 500 compute-sanitizer --tool synccheck --error-exitcode 3 python -m pytest tests/test_gpu_backward.py -m gpu -x -q -k "tf-64-16x16" 2>&1 | tail -4 | sed 's/^/[synccheck tc backward] /' | tee -a $LOG
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'iaf_wg_kernel|iaf_ly_kernel' -s 12 -c 4 -f \
  -o gpurun_out/bwd_tc_c2a python tools/bench_bwd.py c2a 1 > gpurun_out/bwd_tc_c2a.log 2>&1
exit 0
