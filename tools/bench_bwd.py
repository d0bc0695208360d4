"""Time the backward of the IAF step on one GPU: CUDA events around K calls, inputs resident in HBM.  Prints
one JSON line per workload: `bwd_inputs_only` / `bwd_full` = iaf_step_bwd (recomputes the activations), `fwd_plain` /
`fwd_train` = the forward without / with kept activations, `bwd_saved_full` = iaf_step_bwd_saved (what the autograd node runs).
Algorithmic flops of the backward = 3x the forward's live MACs (recompute + data gradient + weight gradient); `algorithmic_tflops`
is that over `bwd_full`.  Tensor-core plans run the backward on the tensor cores (IAF_BWD_TC=0: exact-fp32 SIMT kernels).
usage: python tools/bench_bwd.py [c2a|c2b] [steps]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iaf_b200 import IAFOperator  # noqa: E402
from oracle import iaf_oracle as O  # noqa: E402  (synthetic parameter / input generator only)

WL = {"c2a": [64], "c2b": [160, 160]}


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2a"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    hidden = WL[wl]
    n_z, H, W, B = 32, 16, 16, 256
    hid, hd = O.make_params("tf", n_z, hidden, [n_z, n_z], seed=1)
    z, ctx = O.make_inputs(B, n_z, hidden[0], H, W, seed=0)
    dev = [tuple(torch.from_numpy(np.ascontiguousarray(l[k])).cuda() for k in "Vgb") for l in hid + hd]
    op = IAFOperator("tf", n_z, hidden, [n_z, n_z], nl="elu").set_weights(dev)
    zg, cg = torch.from_numpy(z).cuda(), torch.from_numpy(ctx).cuda()
    g1 = torch.randn_like(zg)
    gl = torch.randn(B, device="cuda")
    out = {}
    for name, need in (("bwd_inputs_only", False), ("bwd_full", True)):
        for _ in range(3):
            op.step_backward(zg, cg, g1, g1, gl, need_params=need)
        torch.cuda.synchronize()
        l0 = op.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            op.step_backward(zg, cg, g1, g1, gl, need_params=need)
        e1.record()
        torch.cuda.synchronize()
        out[name] = {"ms": e0.elapsed_time(e1) / steps, "launches_per_call": (op.launch_count() - l0) // steps}
    # the training pair the autograd node uses: forward that keeps the activations + backward without recompute
    for name, fn in (("fwd_plain", lambda: op._step_raw(zg, cg)), ("fwd_train", lambda: op._step_train_raw(zg, cg))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name] = {"ms": e0.elapsed_time(e1) / steps}
    zo, ls, _, hs = op._step_train_raw(zg, cg)
    for _ in range(3):
        op._backward("step", zg, cg, op._layers, (g1, g1, gl), True, saved=(zo, ls, hs))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        op._backward("step", zg, cg, op._layers, (g1, g1, gl), True, saved=(zo, ls, hs))
    e1.record()
    torch.cuda.synchronize()
    out["bwd_saved_full"] = {"ms": e0.elapsed_time(e1) / steps}
    fwd_flops = op.algorithmic_flops(B, H, W, "cuda:0")
    full = out["bwd_full"]["ms"] * 1e-3
    print(json.dumps({"workload": wl, "B": B, "steps": steps, **out,
                      "algorithmic_flops_bwd": 3 * fwd_flops, "algorithmic_tflops": 3 * fwd_flops / full / 1e12,
                      "backward_path": op.backward_path(H, W, "cuda:0"),
                      "latent_elems_per_s": B * n_z * H * W / full}))


if __name__ == "__main__":
    main()
