"""Oracle-side `iaf_layer` for iaf_b200.elbo.forward (TEST INFRASTRUCTURE ONLY): the stochastic-layer
block of tf_train.py:56-85 evaluated with oracle/iaf_oracle.py in float64 on the CPU."""
import numpy as np
import torch

from . import iaf_oracle as O


class OracleIAF(object):
    def __init__(self, params, hps):
        self.params, self.hps = params, hps

    def __call__(self, scope, eps, post_mean, post_logsd, prior_mean, prior_logsd, context):
        f = lambda t: t.detach().cpu().numpy().astype(np.float64)
        pre = scope + "/ar_multiconv2d/"
        layer = lambda n: {k: f(self.params[pre + n + "/" + k]) for k in "Vgb"}
        hidden, heads = [layer("layer_0"), layer("layer_1")], [layer("layer_out_0"), layer("layer_out_1")]
        zero = np.zeros_like(f(eps))
        r = O.stochastic_layer_down("tf", f(eps), f(post_mean), f(post_logsd), zero, zero, f(prior_mean), f(prior_logsd),
                                    f(context), np.zeros_like(f(context)), hidden, heads, "elu", kl_min=0.0)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(eps.dtype).to(eps.device)
        return t(r["z"]), t(r["kl"].sum(axis=(2, 3))), t(r["kl_cost"])
