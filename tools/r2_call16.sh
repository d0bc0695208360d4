#!/bin/bash
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call16.log
: > $LOG
T="tests/test_gpu_parity_full.py::test_every_sample_of_the_full_batch_against_fp64_oracle"
for cfg in "X=1" "IAF_FZ_STAGE=0" "IAF_TC_FZ=0"; do
  env $cfg timeout 300 python -m pytest "$T" -m gpu -q 2>&1 | grep -E "samples with|passed|failed|Error" | cut -c1-400 | sed "s/^/[$cfg] /" | tee -a $LOG
done
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | sed 's/^/[gpu tests] /' | tee -a $LOG
one() {  # one <label> <batch> [env...]
  lab=$1; b=$2; shift 2
  env "$@" timeout 200 python bench.py --workload c2a --batch $b --steps 300 --warmup 20 --no-cpu-baseline --no-also --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lab', 'B=$b', 'kernel_us', round(d['roofline']['kernel_us'],2))" | tee -a $LOG
}
one "[staged]" 256 X=1; one "[staged]" 256 X=1; one "[staged]" 32 X=1
one "[gathered]" 256 IAF_FZ_STAGE=0
(cd iaf_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DIAF_FZ_PROBE -shared -Xcompiler -fPIC \
     -o ../lib/libiaf_probe.so iaf_capi.cu iaf_pack.cu iaf_simt.cu iaf_tc.cu iaf_bwd.cu 2>&1 | grep -E "error")
echo "== c2a staged"; timeout 120 python tools/probe_run.py c2a 2>&1 | grep PROBE | tee gpurun_out/r2_probe4.log
echo "== c2b"; timeout 120 python tools/probe_run.py c2b 2>&1 | grep PROBE | tee gpurun_out/r2_probe_c2b.log
rm -f iaf_b200/lib/libiaf_probe.so
exit 0
