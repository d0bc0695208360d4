#!/bin/bash
# Run on the GPU box (under gpurun): launch list + one full ncu capture of the step kernel.
#   bash tools/profile.sh <workload: c2a|c2b> <tag>
# Outputs land in gpurun_out/ (scratch); tools/ncu_to_profiles.py turns them into profiles/.
set -u
WL=${1:-c2a}
TAG=${2:-r1}
mkdir -p gpurun_out
# every launch with its device time (cold-cache, serialised: compare shares, not absolutes)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
  --log-file gpurun_out/launches_${WL}_${TAG}.csv \
  python bench.py --workload ${WL} --steps 10 --warmup 3 --no-graph --no-cpu-baseline --no-also --no-e2e > gpurun_out/launches_${WL}_${TAG}.log 2>&1
# the step kernel, once, full set (skip the warm-up launches of that kernel)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'iaf_(fz|tc|simt|ly)_kernel' -s 5 -c 1 -f \
  -o gpurun_out/prof_${WL}_${TAG} \
  python bench.py --workload ${WL} --steps 10 --warmup 3 --no-graph --no-cpu-baseline --no-also --no-e2e > gpurun_out/prof_${WL}_${TAG}.log 2>&1
ls -la gpurun_out | tail -5
