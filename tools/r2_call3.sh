#!/bin/bash
# Round-2 GPU call 3: MMA microbenchmark with a proper (convergent, elected) issue loop; iaf_fz_kernel with unified loader
# warps: parity, timing, timeline; A/B of the epilogue / loader warp split.
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call3.log
: > $LOG
timeout 120 ./tools/mma_bench 2>&1 | tee gpurun_out/r2_mma_bench2.log | grep time
build() {  # build <out> <extra nvcc flags...>
  out=$1; shift
  (cd iaf_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 "$@" -shared -Xcompiler -fPIC \
     -o ../lib/$out iaf_capi.cu iaf_pack.cu iaf_simt.cu iaf_tc.cu iaf_bwd.cu 2>&1 | grep -E "error")
}
one() {  # one <label> <workload> [env...]
  lab=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --workload $wl --steps 300 --warmup 20 --no-cpu-baseline --no-also --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lab', '$wl', 'kernel_us', round(d['roofline']['kernel_us'],2), 'ms_per_step', round(d['ms_per_step']*1e3,2), 'frac', round(d['roofline']['frac'],3))" | tee -a $LOG
}
tl() {  # tl <tag> <flags...>
  tag=$1; shift
  build libiaf_tl.so -DIAF_TC_TIMELINE "$@"
  timeout 120 python tools/tl_run.py c2a > gpurun_out/r2_tl_$tag.log 2>&1
  rm -f iaf_b200/lib/libiaf_tl.so
}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | sed 's/^/[gpu tests, fz EPI8 LD8] /' | tee -a $LOG
one "[fz EPI8 LD8]" c2a X=1; one "[fz EPI8 LD8]" c2a X=1
one "[fz EPI8 LD8 merged=0]" c2a IAF_TC_MERGED=0
tl fz_e8l8
build libiaf_b200.so -DFZ_EPI=16 -DFZ_LD=4
one "[fz EPI16 LD4]" c2a X=1; one "[fz EPI16 LD4]" c2a X=1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1 | sed 's/^/[fz EPI16 LD4] parity: /' | tee -a $LOG
tl fz_e16l4 -DFZ_EPI=16 -DFZ_LD=4
build libiaf_b200.so -DFZ_EPI=16 -DFZ_LD=8
one "[fz EPI16 LD8]" c2a X=1
build libiaf_b200.so
