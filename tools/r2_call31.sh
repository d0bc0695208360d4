#!/bin/bash
# round-2 call 31: backward with the segmented split-K reduction; refreshed ncu evidence of the final kernels
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call31.log
: > $LOG
timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -x -q 2>&1 | tail -2 | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | tee gpurun_out/r2_bench_bwd_c2a_final.json | cut -c1-600 | tee -a $LOG
timeout 100 python tools/bench_bwd.py c2b 10 2>&1 | tail -1 | tee gpurun_out/r2_bench_bwd_c2b_final.json | cut -c1-600 | tee -a $LOG
bash tools/profile.sh c2a r2f >> $LOG 2>&1
bash tools/profile.sh c2b r2f >> $LOG 2>&1
# backward kernels: launch list of one training pair + full capture of the two contraction kernels
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_bwd_c2a_r2f.csv \
  python tools/bench_bwd.py c2a 1 > gpurun_out/launches_bwd_c2a_r2f.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'iaf_lconv|iaf_bwd_wgrad|iaf_bwd_reduce' -s 30 -c 6 -f \
  -o gpurun_out/bwd_c2a_r2f python tools/bench_bwd.py c2a 1 > gpurun_out/bwd_c2a_r2f.log 2>&1
ls -la gpurun_out/*r2f* | tee -a $LOG
exit 0
