#!/bin/bash
# round-2 call 34: weight gradient on the tensor cores (MN-major operands) -- backward tests, training-pair timing
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call34.log
: > $LOG
timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -x -q 2>&1 | tail -25 | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | cut -c1-600 | sed 's/^/[tc dgrad + wgrad] /' | tee -a $LOG
IAF_BWD_WG_TC=0 timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | cut -c1-600 | sed 's/^/[tc dgrad, simt wgrad] /' | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2b 10 2>&1 | tail -1 | cut -c1-600 | sed 's/^/[tc dgrad + wgrad] /' | tee -a $LOG
exit 0
