"""Oracle-side `iaf_layer` callables for iaf_b200.elbo.forward / iaf_b200.elbo_theano.forward (TEST INFRASTRUCTURE
ONLY): the stochastic-layer block of tf_train.py:56-85 / models.py:273-298 evaluated with oracle/iaf_oracle.py in
float64 on the CPU."""
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


class TorchIAF(object):
    """Differentiable oracle iaf_layer: the same block with oracle/iaf_oracle_torch.py ops on the parameters' own
    dtype/device (float64 CPU in the tests), so torch autograd gives the reference gradient of the training objective
    -- what tf.gradients derives in tf_train.py:222-232."""

    def __init__(self, params, hps):
        self.params, self.hps = params, hps

    def __call__(self, scope, eps, post_mean, post_logsd, prior_mean, prior_logsd, context):
        from . import iaf_oracle_torch as OT
        from iaf_b200.elbo import stochastic_layer   # plain torch arithmetic shared with the product's training wrapper
        pre = scope + "/ar_multiconv2d/"
        layer = lambda n: {k: self.params[pre + n + "/" + k] for k in "Vgb"}
        hidden, heads = [layer("layer_0"), layer("layer_1")], [layer("layer_out_0"), layer("layer_out_1")]
        return stochastic_layer(lambda z, c: OT.iaf_step("tf", z, c, hidden, heads, "elu")[:2], eps, post_mean, post_logsd,
                                prior_mean, prior_logsd, context)


class OracleIAFTheano(object):
    """Theano front-end (models.py:273-298): parameters ``{name}_posterior_conv1_{k}_{w,s,b}`` / ``..._out_{k}_...``."""

    def __init__(self, w, hps):
        self.w, self.hps = w, hps

    def __call__(self, name, eps, post_mean, post_logsd, prior_mean, prior_logsd, context):
        f = lambda t: t.detach().cpu().numpy().astype(np.float64)
        pre = name + "_posterior_conv1_"
        layer = lambda n: {k: f(self.w[pre + n + "_" + k]) for k in "wsb"}
        hidden = [layer("%d" % k) for k in range(self.hps["depth_ar"])]
        heads = [layer("out_0"), layer("out_1")]
        zero = np.zeros_like(f(eps))
        r = O.stochastic_layer_down("theano", f(eps), f(post_mean), f(post_logsd), zero, zero, f(prior_mean),
                                    f(prior_logsd), f(context), np.zeros_like(f(context)), hidden, heads,
                                    self.hps["nl"], kl_min=0.0)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(eps.dtype).to(eps.device)
        return t(r["z"]), t(r["kl"].sum(axis=(2, 3))), t(r["kl_cost"])

    def step(self, name, z, context):
        """The bare step (models.py:170-173) for up_iaf2_nl -> (z', arw_logsd)."""
        f = lambda t: t.detach().cpu().numpy().astype(np.float64)
        pre = name + "_posterior_conv1_"
        layer = lambda n: {k: f(self.w[pre + n + "_" + k]) for k in "wsb"}
        hidden = [layer("%d" % k) for k in range(self.hps["depth_ar"])]
        z_new, arw_logsd, _ = O.iaf_step("theano", f(z), f(context), hidden, [layer("out_0"), layer("out_1")], self.hps["nl"])
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(z.dtype).to(z.device)
        return t(z_new), t(arw_logsd)


class TorchIAFTheano(object):
    """Differentiable oracle iaf_layer for the Theano front-end (models.py:273-298 / 169-178) with
    oracle/iaf_oracle_torch.py ops on the parameters' own dtype (float64 CPU in the tests): torch autograd through it is
    the reference for d(cost)/d(parameters), i.e. what ``T.grad`` derives in graphy/misc/optim.py:99-123."""

    def __init__(self, w, hps):
        self.w, self.hps = w, hps

    def _layers(self, name):
        pre = name + "_posterior_conv1_"
        layer = lambda n: {k: self.w[pre + n + "_" + k] for k in "wsb"}
        return [layer("%d" % k) for k in range(self.hps["depth_ar"])], [layer("out_0"), layer("out_1")]

    def step(self, name, z, context):
        from . import iaf_oracle_torch as OT
        hidden, heads = self._layers(name)
        return OT.iaf_step("theano", z, context, hidden, heads, self.hps["nl"])[:2]

    def __call__(self, name, eps, post_mean, post_logsd, prior_mean, prior_logsd, context):
        from iaf_b200.elbo import stochastic_layer   # plain torch arithmetic shared with the product's training wrapper
        return stochastic_layer(lambda z, c: self.step(name, z, c), eps, post_mean, post_logsd, prior_mean, prior_logsd,
                                context)
