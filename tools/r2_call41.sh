#!/bin/bash
# round-2 call 41: fused step prologue of the tensor-core backward (A/B), backward tests
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call41.log
: > $LOG
timeout 400 python -m pytest tests/test_gpu_backward.py -m gpu -x -q 2>&1 | tail -25 | tee -a $LOG
IAF_BWD_FUSED_PROLOGUE=0 timeout 400 python -m pytest tests/test_gpu_backward.py -m gpu -x -q 2>&1 | tail -3 | sed 's/^/[unfused] /' | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | cut -c1-600 | sed 's/^/[fused prologue] /' | tee -a $LOG
IAF_BWD_FUSED_PROLOGUE=0 timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | cut -c1-600 | sed 's/^/[unfused] /' | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2b 10 2>&1 | tail -1 | cut -c1-600 | sed 's/^/[fused prologue] /' | tee -a $LOG
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_backward.py -m gpu -x -q -k "tf-64-16x16 or theano-64-8x8" 2>&1 | tail -3 | sed 's/^/[memcheck] /' | tee -a $LOG
exit 0
