#!/bin/bash
# Round-2 GPU call 1 (under gpurun): confirm the tree, first GPU run of the two opt-in autograd paths, A/B of the
# development flags of the fused kernel, an in-kernel timeline of the default build.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call1.log
: > $LOG
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | sed 's/^/[gpu tests] /' | tee -a $LOG
IAF_MULTICONV_SAVED=1 timeout 200 python -m pytest tests/test_gpu_backward.py -m gpu -q -k "multiconv or factory" 2>&1 | tail -1 | \
  sed 's/^/[IAF_MULTICONV_SAVED=1] /' | tee -a $LOG
IAF_LAYER_AUTOGRAD=1 timeout 200 python - <<'PY' 2>&1 | tail -2 | sed 's/^/[IAF_LAYER_AUTOGRAD=1] /' | tee -a $LOG
import numpy as np, torch, sys
sys.path.insert(0, ".")
from iaf_b200 import elbo
from tests.test_elbo import _setup
from oracle.elbo_oracle import TorchIAF
hps = dict(z_size=32, h_size=64, depth=2, num_blocks=1, kl_min=0.25, image_size=32)
pg, xg, ng = _setup(hps, 3, 9, torch.float32, "cuda"); pc, xc, nc = _setup(hps, 3, 9, torch.float64, "cpu")
for p in (pg, pc):
    for v in p.values(): v.requires_grad_(True)
got = elbo.forward(pg, xg, ng, elbo.CudaIAFTrain(pg, hps, fused=True), hps); ref = elbo.forward(pc, xc, nc, TorchIAF(pc, hps), hps)
got["obj"].backward(); ref["obj"].backward()
worst = max(float((pg[k].grad.double().cpu() - pc[k].grad).abs().max()) / max(float(pc[k].grad.abs().max()), 1e-12) for k in pc if pc[k].grad is not None)
print("fused-layer training gradients: worst relative error %.2e (expect < 2e-3)" % worst)
PY
build() {  # build <out> <extra nvcc flags...>
  out=$1; shift
  (cd iaf_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 "$@" -shared -Xcompiler -fPIC \
     -o ../lib/$out iaf_capi.cu iaf_pack.cu iaf_simt.cu iaf_tc.cu iaf_bwd.cu 2>&1 | grep -E "error")
}
one() {  # one <label> <workload>
  timeout 150 python bench.py --workload $2 --steps 300 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', '$2', round(d['roofline']['kernel_us'],2))" | tee -a $LOG
}
cp iaf_b200/lib/libiaf_b200.so gpurun_out/libiaf_default.so
for variant in "" "-DTC_FAST_EPI" "-DTC_HALO_TRIM" "-DTC_FAST_EPI -DTC_HALO_TRIM" "-DTC_WORKERS=12 -DTC_FAST_EPI -DTC_HALO_TRIM"; do
  build libiaf_b200.so $variant
  one "[$variant]" c2a; one "[$variant]" c2a
  if [ "$variant" = "-DTC_FAST_EPI" ]; then one "[$variant]" c2b; fi
  if [ "$variant" = "" ]; then one "[$variant]" c2b; fi
  if [ "$variant" = "-DTC_FAST_EPI -DTC_HALO_TRIM" ]; then
    timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1 | sed "s/^/[$variant] parity: /" | tee -a $LOG
  fi
done
build libiaf_b200.so   # back to the default build
# in-kernel timeline of CTA 0 (default flags)
build libiaf_tl.so -DIAF_TC_TIMELINE
timeout 120 python tools/tl_run.py c2a > gpurun_out/r2_tl_c2a.log 2>&1
tail -5 gpurun_out/r2_tl_c2a.log
rm -f iaf_b200/lib/libiaf_tl.so
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee -a $LOG
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)|Thread|Core" | tee -a $LOG
nvidia-smi topo -m 2>&1 | head -20 | tee -a $LOG
