"""Pin the oracle (oracle/iaf_oracle.py) against fixtures produced by executing the
reference's own source (tests/golden/make_golden.py), and against the known answers
SURVEY.md section B derives from the reference (mask nnz counts)."""
import os

import numpy as np
import pytest

from oracle import iaf_oracle as O
from tests.golden.cases import MULTICONV_CASES, case_inputs, checksum

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(G, name))


# ---- masks -----------------------------------------------------------------------
SURVEY_NNZ = {  # (n_in, n_out, zerodiag) -> nnz of the 3x3 mask
    (32, 64, False): 9248, (64, 64, False): 18464, (64, 32, True): 9184,
    (32, 160, False): 23120, (160, 160, False): 115280, (160, 32, True): 22960,
}


@pytest.mark.parametrize("key,nnz", sorted(SURVEY_NNZ.items()))
def test_mask_nnz_known_answers(key, nnz):
    n_in, n_out, zd = key
    assert int(O.get_conv_ar_mask(3, 3, n_in, n_out, zd).sum()) == nnz
    # the Theano mask has the same live entries plus the pad channel's 4 non-centre taps
    assert int(O.theano_conv_ar_mask(n_in, n_out, (3, 3), zd).sum()) == nnz + 4 * n_out


def test_masks_match_reference_execution():
    g = load("masks.npz")
    n = 0
    for k in g.files:
        if k.endswith("_nnz"):
            continue
        kind, n_in, n_out, zd = k.split("_")
        n_in, n_out, zd = int(n_in), int(n_out), bool(int(zd))
        if kind == "tf":
            m = O.get_conv_ar_mask(3, 3, n_in, n_out, zd)
        elif kind == "lin":
            m = O.get_linear_ar_mask(n_in, n_out, zd)
        else:
            m = O.theano_conv_ar_mask(n_in, n_out, (3, 3), zd)
        ref = np.unpackbits(g[k])[: m.size].reshape(m.shape)
        assert np.array_equal(ref, m.astype(np.uint8)), k
        n += 1
    assert n >= 40


def test_tf_and_theano_masks_agree_up_to_layout():
    for n_in, n_out in [(4, 8), (8, 4), (32, 64), (64, 32)]:
        for zd in (False, True):
            t = O.get_conv_ar_mask(3, 3, n_in, n_out, zd)            # [ky,kx,ci,co]
            h = O.theano_conv_ar_mask(n_in, n_out, (3, 3), zd)       # [co,ci+1,ky,kx]
            assert np.array_equal(t.transpose(3, 2, 0, 1), h[:, :n_in])
            assert np.all(h[:, n_in, 1, 1] == 0)                     # pad channel never sees the centre


def test_pad2dwithchannel():
    g = load("pad.npz")
    assert np.array_equal(O.pad2dwithchannel(g["x"].astype(np.float64)), g["y"])


# ---- multiconv ----------------------------------------------------------------------
@pytest.mark.parametrize("ci", range(len(MULTICONV_CASES)))
def test_multiconv_matches_reference_execution(ci):
    name, variant, B, n_z, hidden, H, W, nl = MULTICONV_CASES[ci]
    g = load("multiconv.npz")
    hid, heads, z, ctx = case_inputs(variant, B, n_z, hidden, H, W, seed=ci)
    assert checksum(z, ctx, *[v for l in hid + heads for v in l.values()]) == float(g[name + "_insum"])
    f64 = lambda ls: O.cast_params(ls, np.float64)
    m, s = O.multiconv(variant, z.astype(np.float64), ctx.astype(np.float64), f64(hid), f64(heads), nl)
    np.testing.assert_allclose(m, g[name + "_m"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(s, g[name + "_s"], rtol=0, atol=1e-11)


def test_fp32_port_close_to_fp64_truth():
    name, variant, B, n_z, hidden, H, W, nl = MULTICONV_CASES[1]
    g = load("multiconv.npz")
    hid, heads, z, ctx = case_inputs(variant, B, n_z, hidden, H, W, seed=1)
    m, s = O.multiconv(variant, z, ctx, hid, heads, nl)
    assert m.dtype == np.float32
    np.testing.assert_allclose(m, g[name + "_m"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(s, g[name + "_s"], rtol=0, atol=2e-5)


# ---- the AR property: Jacobian triangular, logdet = -sum(0.1 s) -----------------------
@pytest.mark.parametrize("variant", ["tf", "theano"])
def test_jacobian_is_triangular_and_logdet_matches(variant):
    n_z, hidden, H, W = 2, [4], 3, 3
    hid, heads = O.make_params(variant, n_z, hidden, [n_z, n_z], seed=3)
    hid, heads = O.cast_params(hid, np.float64), O.cast_params(heads, np.float64)
    z, ctx = O.make_inputs(1, n_z, hidden[0], H, W, seed=4, dtype=np.float64)
    z1, logsd, logdet = O.iaf_step(variant, z, ctx, hid, heads)
    n = z.size
    J = np.zeros((n, n))
    h = 1e-6
    for j in range(n):
        dz = np.zeros(n); dz[j] = h
        zp, _, _ = O.iaf_step(variant, z + dz.reshape(z.shape), ctx, hid, heads)
        zm, _, _ = O.iaf_step(variant, z - dz.reshape(z.shape), ctx, hid, heads)
        J[:, j] = ((zp - zm) / (2 * h)).reshape(-1)
    # AR order: tf = reverse raster over (y,x) then channel; theano = raster (SURVEY A)
    c, y, x = np.meshgrid(np.arange(n_z), np.arange(H), np.arange(W), indexing="ij")
    pix = (y * W + x).reshape(-1)
    rank = (-pix if variant == "tf" else pix) * n_z + c.reshape(-1)
    order = np.argsort(rank)
    Jo = J[np.ix_(order, order)]
    assert np.max(np.abs(np.triu(Jo, 1))) < 1e-7          # depends only on earlier elements
    np.testing.assert_allclose(np.diag(Jo), np.exp(-logsd).reshape(-1)[order], rtol=1e-6)
    sign, ld = np.linalg.slogdet(J)
    assert sign > 0
    np.testing.assert_allclose(ld, logdet[0], rtol=1e-6)


# ---- IAFLayer.down ----------------------------------------------------------------------
@pytest.mark.parametrize("name", ["kl0", "kl01", "kl5"])
def test_stochastic_layer_down_matches_reference_execution(name):
    g = load("iaflayer_down.npz")
    v = lambda k: g[name + "_" + k]
    hid, heads = O.make_params("tf", 4, [8, 8], [4, 4], seed=77)
    hid, heads = O.cast_params(hid, np.float64), O.cast_params(heads, np.float64)
    r = O.stochastic_layer_down("tf", v("eps"), v("qz_mean"), v("qz_logsd"), v("rz_mean"), v("rz_logsd"),
                                v("pz_mean"), v("pz_logsd"), v("up_context"), v("down_context"),
                                hid, heads, "elu", kl_min=float(v("kl_min")))
    np.testing.assert_allclose(r["z0"], v("z0"), atol=1e-12)
    np.testing.assert_allclose(r["arw_logsd"], 0.1 * v("s"), atol=1e-12)
    np.testing.assert_allclose(r["kl_cost"], v("kl_cost"), atol=1e-9)
    np.testing.assert_allclose(r["kl_obj"], v("kl_obj"), atol=1e-9)


# ---- distributions.py: golden + the reference's own four unit tests -----------------------
def test_distributions_match_reference_execution():
    g = load("distributions.npz")
    a, b = g["a"], g["b"]
    np.testing.assert_allclose(O.logsumexp(a), g["logsumexp"], atol=1e-13)
    np.testing.assert_allclose(O.compute_lowerbound(a.reshape(-1), b.reshape(-1), 4), g["lb_k4"], atol=1e-13)
    np.testing.assert_allclose(O.compute_lowerbound(a.reshape(-1), b.reshape(-1), 1), g["lb_k1"], atol=1e-13)
    assert np.array_equal(O.repeat(a, 3), g["repeat3"])
    np.testing.assert_allclose(O.gaussian_diag_logps(a, 0.3 * b, b), g["logps"], atol=1e-13)


def test_logsumexp():  # tf_utils/distributions_test.py:7-13
    a = np.arange(10)
    res = np.log(np.sum(np.exp(a)))
    np.testing.assert_allclose(O.logsumexp(a.astype(np.float32).reshape([1, -1]))[0], res, rtol=1e-6)


def test_lowerbound():  # tf_utils/distributions_test.py:15-22
    a = np.log(np.array([0.3, 0.3, 0.3, 0.3], np.float32).reshape([1, -1]))
    b = np.log(np.array([0.1, 0.5, 0.9, 0.6], np.float32).reshape([1, -1]))
    res = -(-np.log(4) + np.log(np.sum(np.exp(a - b))))
    assert abs(np.sum(O.compute_lowerbound(a, b, 4)) - res) < 1e-4


def test_lowerbound2():  # tf_utils/distributions_test.py:24-31
    a = np.log(np.array([0.3, 0.3, 0.3, 0.3], np.float32).reshape([-1, 1]))
    b = np.log(np.array([0.1, 0.5, 0.9, 0.6], np.float32).reshape([-1, 1]))
    res = (b - a).sum()
    assert abs(np.sum(O.compute_lowerbound(a, b, 1)) - res) < 1e-4


def test_repeat():  # tf_utils/distributions_test.py:33-38
    a = np.random.RandomState(0).randn(10, 5, 2)
    np.testing.assert_allclose(O.repeat(a, 2), np.repeat(a, 2, axis=0))
