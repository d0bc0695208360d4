#!/bin/bash
# round-2 call 37: new tensor-core-backward tests
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call37.log
: > $LOG
timeout 400 python -m pytest tests/test_gpu_backward.py -m gpu -x -q 2>&1 | tail -25 | tee -a $LOG
exit 0
