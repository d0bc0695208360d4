"""Theano front-end ELBO: the restated `cvae_layer` / `cvae1.f_encode_decode` plumbing is
pinned against vectors produced by executing the reference's own models.py (tests/golden/make_golden_theano_layer.py),
and bits/dim computed with the B200 operator equals bits/dim computed with the oracle operator."""
import os

import numpy as np
import pytest
import torch

from iaf_b200 import elbo_theano as ET
from oracle.elbo_oracle import OracleIAFTheano

GOLD = os.path.join(os.path.dirname(__file__), "golden", "cvae_layer_down.npz")


@pytest.mark.parametrize("name", ["0_1", "1_0", "u0_1", "u1_0"])
def test_layer_matches_reference_models_py(name):
    """cvae_layer.up / down_q (models.py:133-328) for down_iaf2_nl and up_iaf2_nl ("u" cases) + diag prior, incl.
    conv.py's weight-normed convs with pad channel, stride-2 / depth-to-space resampling, nearest-neighbour skip paths,
    and the IAF step."""
    g = np.load(GOLD)
    pre = name + "/"
    w = {k[len(pre) + 2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(pre + "w/")}
    hps = dict(n_z=4, n_h1=8, n_h2=8, depths=[2, 2], depth_ar=1, nl="elu", kl_min=0.0, image_size=16,
               posterior="up_iaf2_nl" if name[0] == "u" else "down_iaf2_nl")
    ds = bool(g[pre + "downsample"])
    up_out, up_state = ET.layer_up(w, name, torch.from_numpy(g[pre + "up_in"]), hps, ds, torch.from_numpy(g[pre + "eps"]),
                                   OracleIAFTheano(w, hps))
    np.testing.assert_allclose(up_out.numpy(), g[pre + "up_out"], rtol=1e-10, atol=1e-10)
    out, kl_bc, kl_sum = ET.layer_down_q(w, name, torch.from_numpy(g[pre + "down_in"]), up_state,
                                         torch.from_numpy(g[pre + "eps"]), OracleIAFTheano(w, hps), hps, ds)
    np.testing.assert_allclose(out.numpy(), g[pre + "down_out"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(kl_bc.numpy(), g[pre + "kl"].sum(axis=(2, 3)), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(kl_sum.numpy(), g[pre + "kl"].sum(axis=(1, 2, 3)), rtol=1e-9, atol=1e-9)


def _setup(hps, B, seed, dtype, device):
    w = {k: torch.from_numpy(np.asarray(v)).to(dtype).to(device) for k, v in ET.make_params(hps, seed=seed).items()}
    rng = np.random.RandomState(seed + 1)
    S = hps["image_size"]
    x = torch.from_numpy(rng.randint(0, 256, size=(B, 3, S, S)).astype(np.uint8)).to(device)
    noise = {}
    for i in range(len(hps["depths"])):
        s = S // 2 ** (i + 1)
        for j in range(hps["depths"][i]):
            noise[(i, j)] = torch.from_numpy(rng.randn(B, hps["n_z"], s, s)).to(dtype).to(device)
    return w, x, noise


@pytest.mark.parametrize("posterior", ["down_iaf2_nl", "up_iaf2_nl"])
def test_forward_free_bits_and_shapes_cpu(posterior):
    hps = dict(n_z=4, n_h1=8, n_h2=8, depths=[2, 2], depth_ar=1, nl="elu", kl_min=0.25, image_size=16, posterior=posterior)
    w, x, noise = _setup(hps, 3, 5, torch.float64, "cpu")
    r = ET.forward(w, x, noise, OracleIAFTheano(w, hps), hps)
    assert r["cost"].shape == (3,) and np.isfinite(float(r["bits_per_dim"]))
    assert sorted(k for k in r if k.startswith("cost_z")) == ["cost_z000_000", "cost_z000_001", "cost_z001_000", "cost_z001_001"]
    # kl_min = 0: the objective is the plain per-sample sum (models.py:465-466); with free bits it is >= that
    hps0 = dict(hps, kl_min=0.0)
    r0 = ET.forward(w, x, noise, OracleIAFTheano(w, hps0), hps0)
    kl = sum(r0[k] for k in r0 if k.startswith("cost_z"))
    np.testing.assert_allclose(r0["cost"].numpy(), ((r0["cost_x"] + kl) / (3 * 16 * 16 * np.log(2.0))).numpy(), rtol=1e-12)
    assert float(r["bits_per_dim"]) >= float(r0["bits_per_dim"]) - 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("hps,B", [
    # C1 shapes (README.md:30): n_h 64, depth_ar 1, three levels -> 16x16, 8x8, 4x4 latents; fused tcgen05 kernel
    (dict(n_z=32, n_h1=64, n_h2=64, depths=[2, 2, 2], depth_ar=1, nl="elu", kl_min=0.25, image_size=32), 4),
    # C4 shapes (README.md:55-58): n_h 160, depth_ar 2, two levels; layer-at-a-time tcgen05 kernel
    (dict(n_z=32, n_h1=160, n_h2=160, depths=[2, 2], depth_ar=2, nl="elu", kl_min=0.25, image_size=32), 2),
    # cvae1's default nonlinearity (models.py:384) runs on the exact-fp32 kernel's run-time switch as well
    (dict(n_z=32, n_h1=64, n_h2=64, depths=[1, 1], depth_ar=1, nl="softplus", kl_min=0.0, image_size=32), 2),
    # the bottom-up placement (up_iaf2_nl, models.py:169-178): the bare step, KL assembled in the top-down pass
    (dict(n_z=32, n_h1=64, n_h2=64, depths=[2, 2], depth_ar=1, nl="elu", kl_min=0.25, image_size=32,
          posterior="up_iaf2_nl"), 4),
    (dict(n_z=32, n_h1=160, n_h2=160, depths=[1, 1], depth_ar=2, nl="elu", kl_min=0.0, image_size=32,
          posterior="up_iaf2_nl"), 2),
])
def test_bits_per_dim_parity_theano(hps, B):
    wg, xg, ng = _setup(hps, B, 9, torch.float32, "cuda")
    wc, xc, nc = _setup(hps, B, 9, torch.float64, "cpu")
    got = ET.forward(wg, xg, ng, ET.CudaIAF(wg, hps), hps)
    ref = ET.forward(wc, xc, nc, OracleIAFTheano(wc, hps), hps)
    rel = abs(float(got["bits_per_dim"]) - float(ref["bits_per_dim"])) / abs(float(ref["bits_per_dim"]))
    assert rel < 1e-4, (float(got["bits_per_dim"]), float(ref["bits_per_dim"]))
    for k in ref:
        if k.startswith("cost_z"):
            np.testing.assert_allclose(got[k].cpu().numpy(), ref[k].numpy(), rtol=2e-4, atol=1e-2)
    np.testing.assert_allclose(got["cost"].cpu().numpy(), ref["cost"].numpy(), rtol=1e-4)
