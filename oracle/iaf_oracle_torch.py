"""torch-CPU restatement of the reference's IAF step (TEST / BASELINE INFRASTRUCTURE ONLY).

Same math as oracle/iaf_oracle.py (which is pinned against the reference's own source,
tests/test_oracle_golden.py), expressed with torch CPU ops so that it uses all host
cores the way the reference's framework would: a dense 3x3 convolution per layer
(``F.conv2d`` stands in for cuDNN's dnn_conv / tf.nn.conv2d) and separate elementwise
ops, with the mask / weight-norm recomputed inside every step exactly as the reference's
graph does (tf_utils/layers.py:53-64, graphy/nodes/ar.py:304-329).  This is what
``bench.py`` times as ``cpu_baseline`` (kind "port") and as ``--impl reference``; the
Theano / TF originals cannot run in this image.  Never imported by iaf_b200/.
"""
import torch
import torch.nn.functional as F

from . import iaf_oracle as O


def _nl(which):
    return {None: lambda h: h, "none": lambda h: h, "elu": F.elu, "softplus": F.softplus, "relu": F.relu,
            "tanh": torch.tanh, "leakyrelu": lambda h: F.leaky_relu(h, 0.01)}[which]


def to_torch(layers, dtype=torch.float32):
    return [{k: torch.from_numpy(v).to(dtype) for k, v in l.items()} for l in layers]


def tf_ar_conv2d(x, layer, zerodiagonal):
    """tf_utils/layers.py:144-154 -> 52-64."""
    V, g, b = layer["V"], layer["g"], layer["b"]
    mask = torch.from_numpy(O.get_conv_ar_mask(3, 3, V.shape[2], V.shape[3], zerodiagonal)).to(V.dtype)
    v = mask * V
    w = torch.exp(g).reshape(1, 1, 1, -1) * v * torch.rsqrt(torch.clamp((v * v).sum(dim=(0, 1, 2), keepdim=True), min=1e-12))
    return F.conv2d(x, w.permute(3, 2, 0, 1), padding=1) + b.reshape(1, -1, 1, 1)


def theano_ar_conv2d(h, layer, zerodiagonal):
    """graphy/nodes/ar.py:304-329 with conv.py:71-83."""
    w, s, b = layer["w"], layer["s"], layer["b"]
    n_out, n_in1 = w.shape[:2]
    mask = torch.from_numpy(O.theano_conv_ar_mask(n_in1 - 1, n_out, (3, 3), zerodiagonal)).to(w.dtype)
    B, C, H, W = h.shape
    hp = torch.zeros((B, C + 1, H + 2, W + 2), dtype=h.dtype)
    hp[:, C] = 1.0
    hp[:, C, 1:-1, 1:-1] = 0.0
    hp[:, :C, 1:-1, 1:-1] = h
    kerns = mask * w
    norm = torch.sqrt((kerns ** 2).sum(dim=(1, 2, 3), keepdim=True)) + 1e-8
    kerns = kerns * (1.0 / norm) * torch.exp(3.0 * s).reshape(-1, 1, 1, 1)
    return F.conv2d(hp, torch.flip(kerns, dims=(2, 3))) + b.reshape(1, -1, 1, 1)


def multiconv(variant, z, context, hidden, heads, nl="elu"):
    f = _nl(nl)
    conv = tf_ar_conv2d if variant == "tf" else theano_ar_conv2d
    x = z
    for i, layer in enumerate(hidden):
        x = conv(x, layer, False)
        if i == 0:
            x = x + context
        x = f(x)
    return [conv(x, layer, True) for layer in heads]


def iaf_step(variant, z, context, hidden, heads, nl="elu", scale=0.1):
    """models.py:281-285 / tf_train.py:69-72."""
    m, s = multiconv(variant, z, context, hidden, heads, nl)
    arw_mean = m * scale
    arw_logsd = s * scale
    z_new = (z - arw_mean) / torch.exp(arw_logsd)
    logdet = -arw_logsd.flatten(1).sum(dim=1)
    return z_new, arw_logsd, logdet
