#!/bin/bash
# round-2 call 35: TC backward with the segmented bias sums; launch list of one training pair
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call35.log
: > $LOG
timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -x -q 2>&1 | tail -5 | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | cut -c1-600 | sed 's/^/[tc dgrad + wgrad] /' | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2b 10 2>&1 | tail -1 | cut -c1-600 | sed 's/^/[tc dgrad + wgrad] /' | tee -a $LOG
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_bwd_c2a_tc.csv \
  python tools/bench_bwd.py c2a 1 > gpurun_out/launches_bwd_c2a_tc.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches_bwd_c2b_tc.csv \
  python tools/bench_bwd.py c2b 1 > gpurun_out/launches_bwd_c2b_tc.log 2>&1
exit 0
