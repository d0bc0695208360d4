"""The SIMT kernels' logic under host emulation of CUDA (tests/emu/cuda_emu.h) against the oracle -- CPU only.

The build container has no GPU; these tests compile iaf_capi.cu / iaf_pack.cu / iaf_simt.cu / iaf_bwd.cu with g++
against an emulation of the CUDA subset they use and drive the SAME C ABI with numpy buffers, so the index arithmetic,
shared-memory staging, barriers and fixed-order reductions of the forward SIMT kernel and of every backward kernel are
executed and checked before GPU time is spent.  The `-m gpu` tests (tests/test_gpu_parity.py) repeat the same
comparisons on the device through libiaf_b200.so; the emulated library is test infrastructure only and cannot be
loaded by iaf_b200/.
Reference for the gradients: torch autograd (fp64) through oracle/iaf_oracle_torch.py, i.e. what theano.grad /
tf.gradients derive for models.py:281-285 + ar.py:396-416 | tf_train.py:69-72 + layers.py:158-166.
"""
import numpy as np
import pytest
import torch

from oracle import iaf_oracle as O
from oracle import iaf_oracle_torch as OT
from tests.emu.harness import EmuOperator

TOL = 2e-5  # fp32 kernels vs fp64 autograd, relative to the largest entry of each tensor


def _keys(variant):
    return ("V", "g", "b") if variant == "tf" else ("w", "s", "b")


def _setup(variant, n_z, hidden, heads, H, W, B, nl):
    hid, hd = O.make_params(variant, n_z, hidden, heads, seed=1)
    z, ctx = O.make_inputs(B, n_z, hidden[0] if hidden else 1, H, W, seed=0)
    layers = [tuple(l[k] for k in _keys(variant)) for l in hid + hd]
    op = EmuOperator(variant, n_z, hidden, heads, H, W, nl=nl).set_weights(layers)
    return op, hid, hd, z, (ctx if hidden else None)


def _torch_params(hid, hd):
    f64 = lambda ls: O.cast_params(ls, np.float64)
    th, thh = OT.to_torch(f64(hid), torch.float64), OT.to_torch(f64(hd), torch.float64)
    for l in th + thh:
        for k in l:
            l[k].requires_grad_(True)
    return th, thh


def _rel(a, b):
    b = b.detach().numpy() if hasattr(b, "detach") else np.asarray(b)
    assert np.isfinite(a).all()
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


STEP_CASES = [
    # variant, n_z, hidden, H, W, B, nl
    ("tf", 4, [8], 4, 4, 2, "elu"),
    ("theano", 4, [8], 4, 4, 2, "elu"),
    ("tf", 8, [16, 16], 5, 7, 2, "elu"),          # two hidden layers, non-square
    ("theano", 8, [16, 16], 6, 9, 1, "softplus"),  # W > 8: two pixel segments; cvae1's default nl (models.py:384)
    ("tf", 6, [12], 3, 5, 2, "tanh"),              # channel counts off the vector widths (padded head / hidden columns)
    ("theano", 6, [12], 3, 5, 2, "relu"),
    ("theano", 4, [], 4, 4, 2, "elu"),             # depth_ar = 0: no hidden layer, context unused
    ("tf", 16, [80], 4, 4, 1, "elu"),              # > 64 channels: several ci / column blocks in the weight gradient
    ("tf", 4, [8], 12, 24, 1, "leakyrelu"),        # several row bands in lconv and in the weight gradient
    ("theano", 4, [4], 2, 2, 40, "elu"),           # more (sample, band) units than weight-gradient CTAs
    ("tf", 32, [64], 16, 16, 2, "elu"),            # C2a shape: the forward kernel splits the image into two row bands
    #                                                (cross-band halo recompute, band partial sums, arrival counters)
]


@pytest.mark.parametrize("variant,n_z,hidden,H,W,B,nl", STEP_CASES)
def test_emulated_step_forward_and_backward(variant, n_z, hidden, H, W, B, nl):
    op, hid, hd, z, ctx = _setup(variant, n_z, hidden, [n_z, n_z], H, W, B, nl)
    # forward (iaf_simt_kernel)
    zo, ls, ld = op.step(z, ctx)
    th, thh = _torch_params(hid, hd)
    zt = torch.from_numpy(z).double().requires_grad_(True)
    ct = torch.from_numpy(ctx).double().requires_grad_(True) if ctx is not None else None
    zn, lsd, ldt = OT.iaf_step(variant, zt, ct, th, thh, nl=nl)
    assert _rel(zo, zn) < 1e-5 and _rel(ls, lsd) < 1e-5 and _rel(ld, ldt) < 1e-5
    # backward
    rng = np.random.RandomState(5)
    gzo, gls = rng.randn(*z.shape).astype(np.float32), rng.randn(*z.shape).astype(np.float32)
    gld = rng.randn(B).astype(np.float32)
    g_z, g_ctx, gw, gs, gb = op.step_bwd(z, ctx, gzo, gls, gld)
    loss = (zn * torch.from_numpy(gzo)).sum() + (lsd * torch.from_numpy(gls)).sum() + (ldt * torch.from_numpy(gld)).sum()
    loss.backward()
    assert _rel(g_z, zt.grad) < TOL
    if ctx is not None:
        assert _rel(g_ctx, ct.grad) < TOL
    for i, l in enumerate(th + thh):
        for g, k in zip((gw[i], gs[i], gb[i]), _keys(variant)):
            assert _rel(g, l[k].grad) < TOL, (i, k)
    # training pair: the forward kernel keeps the hidden activations, the backward uses them instead of recomputing
    zo2, ls2, ld2, hs = op.step_train(z, ctx)
    assert np.array_equal(zo2, zo) and np.array_equal(ls2, ls) and np.array_equal(ld2, ld)
    assert all(np.isfinite(h).all() for h in hs)
    s_z, s_ctx, sw, ss, sb = op.step_bwd_saved(z, ctx, zo2, ls2, hs, gzo, gls, gld)
    assert _rel(s_z, zt.grad) < TOL and (ctx is None or _rel(s_ctx, ct.grad) < TOL)
    for i, l in enumerate(th + thh):
        for g, k in zip((sw[i], ss[i], sb[i]), _keys(variant)):
            assert _rel(g, l[k].grad) < TOL, (i, k)
    # masked taps get exactly zero (the postup contract, ar.py:369-373)
    for i, l in enumerate(th + thh):
        zd = i >= len(hidden)
        if variant == "tf":
            mask = O.get_conv_ar_mask(3, 3, gw[i].shape[2], gw[i].shape[3], zd)
        else:
            mask = O.theano_conv_ar_mask(gw[i].shape[1] - 1, gw[i].shape[0], (3, 3), zd)
        assert (gw[i][mask == 0] == 0).all()


def test_emulated_backward_optional_inputs_and_determinism():
    op, hid, hd, z, ctx = _setup("tf", 4, [8], [4, 4], 4, 4, 3, "elu")
    rng = np.random.RandomState(7)
    gzo = rng.randn(*z.shape).astype(np.float32)
    a = op.step_bwd(z, ctx, gzo, None, None)            # only z' has a gradient
    b = op.step_bwd(z, ctx, gzo, None, None)
    for x, y in zip([a[0], a[1]] + a[2] + a[3] + a[4], [b[0], b[1]] + b[2] + b[3] + b[4]):
        assert np.array_equal(x, y)                        # fixed-order reductions
    th, thh = _torch_params(hid, hd)
    zt, ct = torch.from_numpy(z).double().requires_grad_(True), torch.from_numpy(ctx).double().requires_grad_(True)
    zn, _, _ = OT.iaf_step("tf", zt, ct, th, thh)
    (zn * torch.from_numpy(gzo)).sum().backward()
    assert _rel(a[0], zt.grad) < TOL and _rel(a[1], ct.grad) < TOL
    # input gradients only: the weight-gradient kernels are skipped, g_z / g_ctx unchanged
    c = op.step_bwd(z, ctx, gzo, None, None, params=False)
    assert np.array_equal(c[0], a[0]) and np.array_equal(c[1], a[1])


@pytest.mark.parametrize("variant,heads", [("tf", [4, 4]), ("theano", [4, 4]), ("theano", [8]), ("theano", [6])])
def test_emulated_multiconv_backward(variant, heads):
    n_z, hidden, H, W, B = 4, [8], 4, 5, 2
    if heads == [6]:
        n_z, hidden = 6, [12]
    op, hid, hd, z, ctx = _setup(variant, n_z, hidden, heads, H, W, B, "elu")
    outs = op.multiconv(z, ctx)
    th, thh = _torch_params(hid, hd)
    zt, ct = torch.from_numpy(z).double().requires_grad_(True), torch.from_numpy(ctx).double().requires_grad_(True)
    ref = OT.multiconv(variant, zt, ct, th, thh)
    rng = np.random.RandomState(3)
    g_outs = [rng.randn(*o.shape).astype(np.float32) for o in outs]
    for o, r in zip(outs, ref):
        assert _rel(o, r) < 1e-5
    sum((r * torch.from_numpy(g)).sum() for r, g in zip(ref, g_outs)).backward()
    outs2, hs = op.multiconv_train(z, ctx)   # the training pair of the un-fused operator
    assert all(np.array_equal(a, b) for a, b in zip(outs, outs2))
    for res in (op.multiconv_bwd(z, ctx, g_outs), op.multiconv_bwd_saved(z, ctx, hs, g_outs)):
        g_z, g_ctx, gw, gs, gb = res
        assert _rel(g_z, zt.grad) < TOL and _rel(g_ctx, ct.grad) < TOL
        for i, l in enumerate(th + thh):
            for g, k in zip((gw[i], gs[i], gb[i]), _keys(variant)):
                assert _rel(g, l[k].grad) < TOL, (i, k)


def test_python_autograd_glue_over_the_emulated_abi(monkeypatch):
    """iaf_b200.ops's autograd nodes (_StepFn / _MulticonvFn: argument order, saved tensors, None handling) exercised
    on CPU tensors by pointing the ctypes binding at the emulated library.  Test-only monkeypatching: the product
    refuses CPU tensors (see tests/test_host_cpu.py)."""
    import contextlib
    import ctypes as C
    from iaf_b200 import _lib as L
    from iaf_b200 import ops
    from tests.emu.harness import emu

    def check_input(t, name, shape=None):
        assert isinstance(t, torch.Tensor) and t.dtype == torch.float32
        if shape is not None:
            assert tuple(t.shape) == tuple(shape)
        return t.contiguous()

    monkeypatch.setattr(L, "lib", emu)
    monkeypatch.setattr(ops, "_check_input", check_input)
    monkeypatch.setattr(ops, "_stream", lambda device: C.c_void_p(0))
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())

    variant, n_z, hidden, H, W, B = "tf", 4, [8], 4, 5, 2
    hid, hd = O.make_params(variant, n_z, hidden, [n_z, n_z], seed=1)
    z, ctx = O.make_inputs(B, n_z, hidden[0], H, W, seed=0)
    dev = [tuple(torch.from_numpy(l[k].copy()).requires_grad_(True) for k in "Vgb") for l in hid + hd]
    op = ops.IAFOperator(variant, n_z, hidden, [n_z, n_z], nl="elu", path="simt").set_weights(dev)
    zg, cg = torch.from_numpy(z).requires_grad_(True), torch.from_numpy(ctx).requires_grad_(True)
    th, thh = _torch_params(hid, hd)
    zt, ct = torch.from_numpy(z).double().requires_grad_(True), torch.from_numpy(ctx).double().requires_grad_(True)

    # step: only z' and logdet are used downstream (the logsd gradient arrives as None)
    zo, ls, ld = op.step(zg, cg)
    (zo.square().sum() + 3.0 * ld.sum()).backward()
    zn, _, ldt = OT.iaf_step(variant, zt, ct, th, thh)
    (zn.square().sum() + 3.0 * ldt.sum()).backward()
    assert _rel(zg.grad.numpy(), zt.grad) < TOL and _rel(cg.grad.numpy(), ct.grad) < TOL
    for i, l in enumerate(th + thh):
        for t, k in zip(dev[i], "Vgb"):
            assert _rel(t.grad.numpy(), l[k].grad) < TOL, (i, k)

    # input gradients only (parameters frozen): the parameter-gradient outputs are skipped
    frozen = [tuple(t.detach() for t in l) for l in dev]
    op2 = ops.IAFOperator(variant, n_z, hidden, [n_z, n_z], nl="elu", path="simt").set_weights(frozen)
    z2 = torch.from_numpy(z).requires_grad_(True)
    op2.step(z2, torch.from_numpy(ctx))[0].square().sum().backward()
    zt2 = torch.from_numpy(z).double().requires_grad_(True)
    OT.iaf_step(variant, zt2, ct.detach(), [{k: v.detach() for k, v in l.items()} for l in th],
                [{k: v.detach() for k, v in l.items()} for l in thh])[0].square().sum().backward()
    assert _rel(z2.grad.numpy(), zt2.grad) < TOL

    # un-fused operator, second head unused (first with the recompute backward, IAF_MULTICONV_SAVED=0)
    monkeypatch.setenv("IAF_MULTICONV_SAVED", "0")
    for t in [zg, cg] + [t for l in dev for t in l]:
        t.grad = None
    m, s = op.multiconv(zg, cg)
    m.sum().backward()
    for t in [zt, ct] + [v for l in th + thh for v in l.values()]:
        t.grad = None
    OT.multiconv(variant, zt, ct, th, thh)[0].sum().backward()
    assert _rel(zg.grad.numpy(), zt.grad) < TOL
    assert _rel(dev[1][0].grad.numpy(), thh[0]["V"].grad) < TOL
    assert float(dev[2][0].grad.abs().max()) == 0.0   # head 1 received no gradient

    # the default kept-activation path of the un-fused operator gives the same gradients
    monkeypatch.delenv("IAF_MULTICONV_SAVED")
    ref_gz, ref_gv = zg.grad.clone(), dev[1][0].grad.clone()
    for t in [zg, cg] + [t for l in dev for t in l]:
        t.grad = None
    op.multiconv(zg, cg)[0].sum().backward()
    assert torch.allclose(zg.grad, ref_gz, rtol=1e-5, atol=1e-6) and torch.allclose(dev[1][0].grad, ref_gv, rtol=1e-5, atol=1e-6)

    # autograd node of the fused layer block (iaf_layer_fwd / iaf_layer_bwd) against the same block built from
    # torch ops around the differentiable step
    from iaf_b200.elbo import stochastic_layer
    g = torch.Generator().manual_seed(4)
    mk = lambda s_=1.0: (s_ * torch.randn(z.shape, generator=g)).requires_grad_(True)
    eps_t, pm, pls, prm, prl = torch.randn(z.shape, generator=g), mk(), mk(0.3), mk(), mk(0.3)
    for t in [cg] + [t for l in dev for t in l]:
        t.grad = None
    z1, _, kl_bc, kl_cost = op.layer(eps_t, pm, pls, prm, prl, cg, want_kl=False)
    (z1.square().sum() + torch.clamp(kl_bc.mean(dim=0), min=0.25).sum() + 0.5 * kl_cost.sum()).backward()
    got = [t.grad.clone() for t in (pm, pls, prm, prl, cg, dev[0][0], dev[2][1])]
    for t in [pm, pls, prm, prl, cg] + [t for l in dev for t in l]:
        t.grad = None
    z2, bc2, cost2 = stochastic_layer(lambda a, b: op.step(a, b, want_logdet=False)[:2], eps_t, pm, pls, prm, prl, cg)
    (z2.square().sum() + torch.clamp(bc2.mean(dim=0), min=0.25).sum() + 0.5 * cost2.sum()).backward()
    for a, t in zip(got, (pm, pls, prm, prl, cg, dev[0][0], dev[2][1])):
        assert float((a - t.grad).abs().max()) <= 2e-5 * max(float(t.grad.abs().max()), 1e-6)

    # the reference driver's NaN guard (graphy/function.py:107-110), opt-in
    opn = ops.IAFOperator(variant, n_z, hidden, [n_z, n_z], nl="elu", path="simt", checknan="raise").set_weights(frozen)
    opn.step(torch.from_numpy(z), torch.from_numpy(ctx))
    zbad = torch.from_numpy(z).clone()
    zbad[1, 0, 0, 0] = float("nan")
    with pytest.raises(FloatingPointError):
        opn.step(zbad, torch.from_numpy(ctx), want_logdet=False)

    # no grad requested: plain call, nothing recorded
    with torch.no_grad():
        assert not op.step(zg, cg)[0].requires_grad


# ---------------------------------------------------------------------------------------
# the forward SIMT kernel's source against the fixtures produced by executing the reference's own code
# ---------------------------------------------------------------------------------------
import os  # noqa: E402

from tests.golden.cases import MULTICONV_CASES, case_inputs  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("ci", range(len(MULTICONV_CASES)), ids=[c[0] for c in MULTICONV_CASES])
def test_emulated_multiconv_against_reference_fixtures(ci):
    """iaf_multiconv_fwd (iaf_pack_kernel + iaf_simt_kernel under emulation) == what the reference's ar_multiconv2d /
    multiconv2d source produced (tests/golden/make_golden.py); the GPU suite repeats this on the device."""
    name, variant, B, n_z, hidden, H, W, nl = MULTICONV_CASES[ci]
    g = np.load(os.path.join(GOLD, "multiconv.npz"))
    hid, heads, z, ctx = case_inputs(variant, B, n_z, hidden, H, W, seed=ci)
    layers = [tuple(l[k] for k in _keys(variant)) for l in hid + heads]
    op = EmuOperator(variant, n_z, hidden, [n_z, n_z], H, W, nl=nl).set_weights(layers)
    m, s = op.multiconv(z, ctx if hidden else None)
    for got, ref in ((m, g[name + "_m"]), (s, g[name + "_s"])):
        assert float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1.0)) < 2e-5


@pytest.mark.parametrize("name", ["kl0", "kl01", "kl5"])
def test_emulated_fused_layer_against_iaflayer_down_fixture(name):
    """iaf_layer_fwd (SIMT kernel, layer mode, under emulation) vs the tensors IAFLayer.down (tf_train.py:46-95, executed
    from the reference source) produced: z', kl_cost and, through the rank-local free-bits rule, kl_obj."""
    g = np.load(os.path.join(GOLD, "iaflayer_down.npz"))
    v = lambda k: g[name + "_" + k]
    hid, heads = O.make_params("tf", 4, [8, 8], [4, 4], seed=77)
    layers = [tuple(l[k] for k in "Vgb") for l in hid + heads]
    H, W = v("eps").shape[2:]
    op = EmuOperator("tf", 4, [8, 8], [4, 4], H, W, nl="elu").set_weights(layers)
    z1, kl, kl_bc, kl_cost = op.layer(v("eps"), v("rz_mean") + v("qz_mean"), v("rz_logsd") + v("qz_logsd"), v("pz_mean"),
                                      v("pz_logsd"), v("up_context") + v("down_context"))
    rel = lambda a, ref: float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1.0))
    assert rel(z1, (v("z0") - 0.1 * v("m")) / np.exp(0.1 * v("s"))) < 1e-5
    assert rel(kl_cost, v("kl_cost")) < 1e-5
    assert rel(kl.sum(axis=(2, 3)), kl_bc) < 1e-5
    kl_min = float(v("kl_min"))
    kl_obj = np.maximum(kl_bc.mean(axis=0, keepdims=True), kl_min).repeat(kl_bc.shape[0], 0).sum(axis=1) if kl_min > 0 else kl_cost
    assert rel(kl_obj, v("kl_obj")) < 1e-5


def test_emulated_random_shapes_forward_and_both_backward_paths():
    """Seeded random configurations (n_z down to 1, widening and narrowing hidden layers, 1 x N and N x 1 maps, every
    nonlinearity, both variants): forward, backward with recompute and backward with kept activations, each against
    fp64 autograd.  (tools-free twin of the ad-hoc stress runs of round 1: three seeds x 14 draws, worst error 1.5e-6.)"""
    rng = np.random.RandomState(1)
    done = 0
    while done < 6:
        variant = str(rng.choice(["tf", "theano"]))
        n_z = int(rng.choice([1, 2, 3, 4, 6, 8]))
        nh = int(rng.choice([0, 1, 2])) if variant == "theano" else int(rng.choice([1, 2]))
        hidden, prev = [], n_z
        for _ in range(nh):
            prev = int(rng.choice([prev * k for k in (1, 2, 3)] + [prev // k for k in (2, 3) if prev % k == 0 and prev // k > 0]))
            hidden.append(prev)
        H, W = int(rng.choice([1, 2, 3, 5, 9, 17])), int(rng.choice([1, 2, 3, 7, 8, 9, 16, 17, 25]))
        B, nl = int(rng.choice([1, 2, 3])), str(rng.choice(["elu", "softplus", "relu", "tanh", "leakyrelu"]))
        if not (prev % n_z == 0 or n_z % prev == 0):
            continue
        done += 1
        op, hid, hd, z, ctx = _setup(variant, n_z, hidden, [n_z, n_z], H, W, B, nl)
        zo, ls, ld, hs = op.step_train(z, ctx)
        th, thh = _torch_params(hid, hd)
        zt = torch.from_numpy(z).double().requires_grad_(True)
        ct = torch.from_numpy(ctx).double().requires_grad_(True) if ctx is not None else None
        zn, lsd, ldt = OT.iaf_step(variant, zt, ct, th, thh, nl=nl)
        g1, g2 = rng.randn(*z.shape).astype(np.float32), rng.randn(*z.shape).astype(np.float32)
        g3 = rng.randn(B).astype(np.float32)
        ((zn * torch.from_numpy(g1)).sum() + (lsd * torch.from_numpy(g2)).sum() + (ldt * torch.from_numpy(g3)).sum()).backward()
        tag = (variant, n_z, hidden, H, W, B, nl)
        assert _rel(zo, zn) < 1e-5 and _rel(ld, ldt) < 1e-5, tag
        for res in (op.step_bwd(z, ctx, g1, g2, g3), op.step_bwd_saved(z, ctx, zo, ls, hs, g1, g2, g3)):
            g_z, g_ctx, gw, gs, gb = res
            assert _rel(g_z, zt.grad) < TOL, tag
            if ctx is not None:
                assert _rel(g_ctx, ct.grad) < TOL, tag
            for i, l in enumerate(th + thh):
                for g, k in zip((gw[i], gs[i], gb[i]), _keys(variant)):
                    if float(l[k].grad.abs().max()) > 1e-9:
                        assert _rel(g, l[k].grad) < TOL, (tag, i, k)


@pytest.mark.parametrize("variant,n_z,hidden,H,W,B,nl,which", [
    ("tf", 4, [8, 8], 6, 6, 3, "elu", "all"),          # IAFLayer.down's shape family (tf_train.py:69: two hidden layers)
    ("theano", 4, [8], 5, 7, 2, "softplus", "all"),    # cvae_layer down_q (models.py:273-298)
    ("tf", 6, [12], 3, 5, 2, "elu", "bc_only"),        # only the free-bits input (kl_bc) carries a gradient
    ("theano", 4, [], 4, 4, 2, "elu", "z_only"),       # depth_ar = 0, only z' used downstream
])
def test_emulated_fused_layer_backward(variant, n_z, hidden, H, W, B, nl, which):
    """iaf_layer_bwd (the elementwise pre / affine / post kernels around the stack's backward) against fp64 autograd
    through the restated block: posterior sample, logqs, the step, prior logps, kl and its two reductions."""
    import math
    op, hid, hd, _, _ = _setup(variant, n_z, hidden, [n_z, n_z], H, W, B, nl)
    rng = np.random.RandomState(11)
    shp = (B, n_z, H, W)
    eps, pm, prm = (rng.randn(*shp).astype(np.float32) for _ in range(3))
    pls, prl = (0.3 * rng.randn(*shp).astype(np.float32) for _ in range(2))
    ctx = (0.1 * rng.randn(B, hidden[0], H, W)).astype(np.float32) if hidden else None
    g_z = rng.randn(*shp).astype(np.float32) if which in ("all", "z_only") else None
    g_kl = rng.randn(*shp).astype(np.float32) if which == "all" else None
    g_bc = rng.randn(B, n_z).astype(np.float32) if which in ("all", "bc_only") else None
    g_cost = rng.randn(B).astype(np.float32) if which == "all" else None
    # forward agrees with the block first
    zo, kl, kl_bc, kl_cost = op.layer(eps, pm, pls, prm, prl, ctx if ctx is not None else np.zeros((B, 1, H, W), np.float32))
    th, thh = _torch_params(hid, hd)
    t = lambda a: torch.from_numpy(a).double().requires_grad_(True)
    te, tpm, tpls, tprm, tprl = t(eps), t(pm), t(pls), t(prm), t(prl)
    tc = t(ctx) if ctx is not None else None
    c = 0.5 * math.log(2.0 * math.pi)
    z0 = tpm + torch.exp(tpls) * te
    zn, lsd, _ = OT.iaf_step(variant, z0, tc, th, thh, nl=nl)
    klt = (-c - tpls - 0.5 * te * te + lsd) - (-c - tprl - 0.5 * (zn - tprm) ** 2 * torch.exp(-2.0 * tprl))
    assert _rel(zo, zn) < 1e-5 and _rel(kl, klt) < 1e-5 and _rel(kl_cost, klt.sum(dim=(1, 2, 3))) < 1e-5
    loss = 0.0
    if g_z is not None:
        loss = loss + (zn * torch.from_numpy(g_z)).sum()
    if g_kl is not None:
        loss = loss + (klt * torch.from_numpy(g_kl)).sum()
    if g_bc is not None:
        loss = loss + (klt.sum(dim=(2, 3)) * torch.from_numpy(g_bc)).sum()
    if g_cost is not None:
        loss = loss + (klt.sum(dim=(1, 2, 3)) * torch.from_numpy(g_cost)).sum()
    loss.backward()
    outs, g_ctx, gw, gs, gb = op.layer_bwd(eps, pm, pls, prm, prl, ctx, g_z, g_kl, g_bc, g_cost)
    for got, ref, name in zip(outs, (tpm, tpls, tprm, tprl, te), ("post_mean", "post_logsd", "prior_mean", "prior_logsd", "eps")):
        r = ref.grad if ref.grad is not None else torch.zeros_like(ref)
        if float(r.abs().max()) > 1e-12:
            assert _rel(got, r) < TOL, name
        else:
            assert float(np.abs(got).max()) < 1e-12, name
    if ctx is not None:
        assert _rel(g_ctx, tc.grad) < TOL
    for i, l in enumerate(th + thh):
        for g, k in zip((gw[i], gs[i], gb[i]), _keys(variant)):
            if float(l[k].grad.abs().max()) > 1e-9:
                assert _rel(g, l[k].grad) < TOL, (i, k)


@pytest.mark.parametrize("san", ["thread", "address,undefined"])
def test_emulated_kernels_under_sanitizers(tmp_path, san):
    """tests/emu/race_check.cc built with a sanitizer.
    thread: with one std::thread per CUDA thread and a std::barrier per __syncthreads, a missing barrier or an
    unsynchronised reuse of shared memory in a kernel IS a data race ThreadSanitizer reports (checked once by hand that
    the detector bites: with __syncthreads compiled out the same binary reports 254 races).
    address,undefined: out-of-range global / shared-memory indices (the dynamic shared memory's slack is poisoned) and
    undefined integer behaviour in the index arithmetic.
    Covers weight packing, the forward SIMT kernel in all three modes, and every backward kernel, both variants."""
    import subprocess
    here = os.path.dirname(__file__)
    root = os.path.dirname(here)
    csrc = os.path.join(root, "iaf_b200", "csrc")
    exe = str(tmp_path / "race_check")
    cmd = ["g++", "-std=c++20", "-O1", "-g", "-fsanitize=" + san, "-fno-sanitize-recover=undefined", "-pthread", "-DIAF_EMU",
           "-w", "-I", os.path.join(here, "emu"), "-I", csrc]
    for f in ("iaf_capi.cu", "iaf_pack.cu", "iaf_simt.cu", "iaf_bwd.cu"):
        cmd += ["-x", "c++", os.path.join(csrc, f)]
    cmd += ["-x", "c++", os.path.join(here, "emu", "tc_stub.cc"), "-x", "c++", os.path.join(here, "emu", "race_check.cc"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and ("san" in (r.stderr + r.stdout).lower() and "cannot find" in (r.stderr + r.stdout).lower()):
        pytest.skip("sanitizer runtime not available: " + r.stderr[-200:])
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=66", ASAN_OPTIONS="detect_leaks=0"))
    assert "Sanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[:3000]
    assert r.returncode == 0 and "race_check ok" in r.stdout, (r.returncode, r.stdout[-500:], r.stderr[-500:])
