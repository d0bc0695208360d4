#!/bin/bash
# Round-2 GPU call 6: where do the 12.5 us of the empty fz kernel go?  (IAF_FZ_DBG=31: barrier hand-offs only)
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call6.log
: > $LOG
one() {  # one <label> <batch> [env...]
  lab=$1; b=$2; shift 2
  env "$@" timeout 200 python bench.py --workload c2a --batch $b --steps 300 --warmup 20 --no-cpu-baseline --no-also --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lab', 'B=$b', 'kernel_us', round(d['roofline']['kernel_us'],2), d['config']['launch'])" | tee -a $LOG
}
for b in 256 128 64 32 8; do one "[dbg31]" $b IAF_FZ_DBG=31; done
for b in 256 32; do one "[dbg63: no weight load]" $b IAF_FZ_DBG=63; done
for b in 256 32; do one "[dbg31 PDL off]" $b IAF_FZ_DBG=31 IAF_PDL=0; done
for b in 256 32; do one "[dbg63 PDL off]" $b IAF_FZ_DBG=63 IAF_PDL=0; done
for b in 256 128 64 32 8; do one "[real]" $b X=1; done
for b in 256 32; do one "[real gen1]" $b IAF_TC_FZ=0; done
for b in 256 32; do one "[real PDL off]" $b IAF_PDL=0; done
# direct launches instead of a graph
timeout 200 python bench.py --workload c2a --steps 300 --warmup 20 --no-cpu-baseline --no-also --no-e2e --no-graph 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[real, direct launches] kernel_us', round(d['roofline']['kernel_us'],2))" | tee -a $LOG
exit 0
