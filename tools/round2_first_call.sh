#!/bin/bash
# First GPU call of the next round (run under gpurun, ~4 min): confirms the tree, then A/Bs the two development variants
# that were written after round 1's GPU budget ran out (tools/experiments/README.md) and takes a fresh capture of the
# backward kernels.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2_gpu_tests.log
# opt-in kept-activation path of the un-fused operator's autograd node (emulation-tested only so far)
IAF_MULTICONV_SAVED=1 timeout 200 python -m pytest tests/test_gpu_backward.py -m gpu -q -k "multiconv or factory" 2>&1 | tail -1 | \
  sed 's/^/[IAF_MULTICONV_SAVED=1] /' | tee -a gpurun_out/r2_ab.log
# opt-in autograd node of the fused layer block (iaf_layer_bwd: emulation-tested only so far)
IAF_LAYER_AUTOGRAD=1 timeout 200 python - <<'PY' 2>&1 | tail -2 | sed 's/^/[IAF_LAYER_AUTOGRAD=1] /' | tee -a gpurun_out/r2_ab.log
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
build() {  # build <extra nvcc flags...>
  (cd iaf_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 "$@" -shared -Xcompiler -fPIC \
     -o ../lib/libiaf_b200.so iaf_capi.cu iaf_pack.cu iaf_simt.cu iaf_tc.cu iaf_bwd.cu 2>&1 | grep -E "error")
}
one() {  # one <label> <workload>
  timeout 150 python bench.py --workload $2 --steps 300 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', '$2', round(d['roofline']['kernel_us'],2))" | tee -a gpurun_out/r2_ab.log
}
# TC_WORKERS=12: 14 warps per CTA are allocated as 16, which lifts the register cap from 96 (18 warps -> 20) to 128
for variant in "" "-DTC_FAST_EPI" "-DTC_HALO_TRIM" "-DTC_FAST_EPI -DTC_HALO_TRIM" "-DTC_WORKERS=12" "-DTC_WORKERS=12 -DTC_FAST_EPI -DTC_HALO_TRIM"; do
  build $variant
  one "[$variant]" c2a; one "[$variant]" c2a
  if [ -n "$variant" ]; then
    timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1 | sed "s/^/[$variant] parity: /" | tee -a gpurun_out/r2_ab.log
  fi
done
build   # back to the default build
timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | sed 's/^/[default] /' | tee -a gpurun_out/r2_ab.log
build -DBW_FASTDIV
timeout 100 python tools/bench_bwd.py c2a 20 2>&1 | tail -1 | sed 's/^/[-DBW_FASTDIV] /' | tee -a gpurun_out/r2_ab.log
timeout 200 python -m pytest tests/test_gpu_backward.py -m gpu -x -q 2>&1 | tail -1 | sed 's/^/[-DBW_FASTDIV] backward parity: /' | tee -a gpurun_out/r2_ab.log
build   # back to the default build
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'iaf_lconv|iaf_bwd_wgrad' -s 30 -c 5 -f \
  -o gpurun_out/r2_bwd_c2a python tools/bench_bwd.py c2a 1 > gpurun_out/r2_ncu_bwd.log 2>&1
# memory / shared-memory hazard checks of the real kernels on one small case each (compute-sanitizer is in the image)
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_step_against_fp64_oracle and tf-64-16x16" 2>&1 | tail -3 | sed 's/^/[memcheck] /' | tee -a gpurun_out/r2_ab.log
timeout 240 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_backward.py -m gpu -x -q -k "tf-8x8-5x7 or 16x16-5x7" 2>&1 | tail -3 | sed 's/^/[racecheck] /' | tee -a gpurun_out/r2_ab.log
