#!/bin/bash
# round-2 call 39: backward tests incl. the batch-shrink case, with the tail-zeroing image kernel
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call39.log
: > $LOG
timeout 400 python -m pytest tests/test_gpu_backward.py -m gpu -x -q 2>&1 | tail -25 | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | cut -c1-600 | sed 's/^/[tc backward] /' | tee -a $LOG
exit 0
