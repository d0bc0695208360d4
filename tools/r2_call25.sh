#!/bin/bash
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call25.log
: > $LOG
T1="tests/test_gpu_parity.py::test_step_against_fp64_oracle"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest $T1 -m gpu -x -q -k "tf-64-16x16 or theano-64-16x16 or theano-64-8x8" 2>&1 | tail -3 | sed 's/^/[memcheck fz] /' | tee -a $LOG
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest $T1 -m gpu -x -q -k "tf-64-16x16" 2>&1 | tail -4 | sed 's/^/[racecheck fz staged] /' | tee -a $LOG
timeout 400 compute-sanitizer --tool synccheck --error-exitcode 3 python -m pytest $T1 -m gpu -x -q -k "tf-64-16x16 or theano-64-8x8" 2>&1 | tail -3 | sed 's/^/[synccheck fz] /' | tee -a $LOG
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity_full.py -m gpu -x -q -k "tensor_core_fused_layer or imported" 2>&1 | tail -3 | sed 's/^/[memcheck layer\/import] /' | tee -a $LOG
exit 0
