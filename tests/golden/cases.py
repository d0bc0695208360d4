"""Case table shared by tests/golden/make_golden.py (which runs the reference's source)
and the tests (which regenerate the same seeded inputs)."""
import numpy as np

from oracle import iaf_oracle as O

MULTICONV_CASES = [
    # name, variant, B, n_z, hidden, H, W, nl
    ("tf_small", "tf", 3, 4, [8, 8], 6, 5, "elu"),
    ("tf_c2a", "tf", 2, 32, [64], 16, 16, "elu"),
    ("tf_c2b", "tf", 1, 32, [160, 160], 16, 16, "elu"),
    ("tf_wide_in", "tf", 2, 8, [8], 4, 4, "elu"),
    ("th_small", "theano", 3, 4, [8], 5, 6, "elu"),
    ("th_c1_16", "theano", 2, 32, [64], 16, 16, "elu"),
    ("th_c1_8", "theano", 2, 32, [64], 8, 8, "elu"),
    ("th_c1_4", "theano", 2, 32, [64], 4, 4, "elu"),
    ("th_c4_8", "theano", 1, 32, [160, 160], 8, 8, "softplus"),
    ("th_depth0", "theano", 2, 4, [], 5, 5, "elu"),
    ("th_relu", "theano", 2, 4, [8, 8], 3, 7, "relu"),
]


def checksum(*arrays):
    return float(sum(np.float64(np.sum(np.asarray(a, dtype=np.float64) * (np.arange(1, a.size + 1).reshape(a.shape) % 7.0)))
                     for a in arrays))


def case_inputs(variant, B, n_z, hidden, H, W, seed):
    hid, heads = O.make_params(variant, n_z, hidden, [n_z, n_z], seed=seed + 100)
    n_ctx = hidden[0] if hidden else n_z
    z, ctx = O.make_inputs(B, n_z, n_ctx, H, W, seed=seed)
    return hid, heads, z, ctx
