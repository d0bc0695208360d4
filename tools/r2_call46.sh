#!/bin/bash
# round-2 call 46: merged MMAs in the data-gradient stages (A/B), backward tests
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call46.log
: > $LOG
timeout 400 python -m pytest tests/test_gpu_backward.py -m gpu -x -q 2>&1 | tail -25 | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | cut -c1-700 | sed 's/^/[merged dgrad] /' | tee -a $LOG
IAF_DG_MERGED=0 timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | cut -c1-700 | sed 's/^/[separate] /' | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2b 10 2>&1 | tail -1 | cut -c1-700 | sed 's/^/[merged dgrad] /' | tee -a $LOG
IAF_DG_MERGED=0 timeout 100 python tools/bench_bwd.py c2b 10 2>&1 | tail -1 | cut -c1-700 | sed 's/^/[separate] /' | tee -a $LOG
exit 0
