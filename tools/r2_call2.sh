#!/bin/bash
# Round-2 GPU call 2: MMA operand-layout microbenchmark; first run of iaf_fz_kernel (parity, A/B against the first-generation
# kernel, in-kernel timeline).
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call2.log
: > $LOG
timeout 120 ./tools/mma_bench 2>&1 | tee gpurun_out/r2_mma_bench.log | tail -45
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | sed 's/^/[gpu tests, fz] /' | tee -a $LOG
one() {  # one <label> <workload> [env...]
  lab=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --workload $wl --steps 300 --warmup 20 --no-cpu-baseline --no-also --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lab', '$wl', 'kernel_us', round(d['roofline']['kernel_us'],2), 'ms_per_step', round(d['ms_per_step']*1e3,2), 'frac', round(d['roofline']['frac'],3))" | tee -a $LOG
}
one "[fz]" c2a X=1; one "[fz]" c2a X=1
one "[fz merged=0]" c2a IAF_TC_MERGED=0
one "[gen1]" c2a IAF_TC_FZ=0
one "[c2b]" c2b X=1
# in-kernel timeline of CTA 0
(cd iaf_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DIAF_TC_TIMELINE -shared -Xcompiler -fPIC \
     -o ../lib/libiaf_tl.so iaf_capi.cu iaf_pack.cu iaf_simt.cu iaf_tc.cu iaf_bwd.cu 2>&1 | grep -E "error")
timeout 120 python tools/tl_run.py c2a > gpurun_out/r2_tl_fz.log 2>&1
tail -3 gpurun_out/r2_tl_fz.log
rm -f iaf_b200/lib/libiaf_tl.so
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_step_against_fp64_oracle and tf-64-16x16" 2>&1 | tail -4 | sed 's/^/[memcheck] /' | tee -a $LOG
