#!/bin/bash
# round-2 call 32: data gradient on the tensor cores -- backward tests, training-pair timing, SIMT for comparison
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call32.log
: > $LOG
timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -x -q 2>&1 | tail -15 | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | cut -c1-600 | sed 's/^/[tc dgrad] /' | tee -a $LOG
IAF_BWD_TC=0 timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | cut -c1-600 | sed 's/^/[simt] /' | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2b 10 2>&1 | tail -1 | cut -c1-600 | sed 's/^/[tc dgrad] /' | tee -a $LOG
exit 0
