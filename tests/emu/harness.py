"""Drive tests/emu/_build/libiaf_emu.so (the C ABI compiled against the host emulation of CUDA) with numpy buffers.
TEST INFRASTRUCTURE ONLY: it exists so that the kernels' logic is exercised by the CPU test-suite."""
import ctypes as C

import numpy as np

from iaf_b200 import _lib as L
from . import build_emu

_emu = None


def emu():
    global _emu
    if _emu is None:
        lib = C.CDLL(build_emu.build())
        for name, (res, args) in L.SYMBOLS.items():
            f = getattr(lib, name)
            f.restype = res
            f.argtypes = args
        _emu = lib
    return _emu


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


def _arr(arrays):
    return (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])


def _check(st):
    if st != 0:
        raise RuntimeError("emu status %d: %s" % (st, emu().iaf_strerror(st).decode()))


class EmuOperator(object):
    """One plan of the emulated library.  layers: list of (w, scale, bias) float32 numpy arrays, reference layouts."""

    def __init__(self, variant, n_z, hidden, heads, H, W, nl="elu"):
        self.lib = emu()
        d = L.IafDesc()
        d.variant = L.VARIANTS[variant]
        d.n_z = n_z
        d.n_hidden = len(hidden)
        for i, h in enumerate(hidden):
            d.hidden[i] = h
        d.n_heads = len(heads)
        for i, h in enumerate(heads):
            d.head[i] = h
        d.H, d.W, d.nl, d.path = H, W, L.NLS[nl], L.PATHS["simt"]
        self.n_z, self.hidden, self.heads, self.H, self.W = n_z, list(hidden), list(heads), H, W
        self.plan = C.c_void_p()
        _check(self.lib.iaf_plan_create(C.byref(self.plan), C.byref(d)))
        self.layers = None

    def __del__(self):
        try:
            self.lib.iaf_plan_destroy(self.plan)
        except Exception:
            pass

    def set_weights(self, layers):
        self.layers = [tuple(np.ascontiguousarray(t, dtype=np.float32) for t in l) for l in layers]
        _check(self.lib.iaf_pack_weights(self.plan, _arr([l[0] for l in self.layers]), _arr([l[1] for l in self.layers]),
                                         _arr([l[2] for l in self.layers]), None))
        return self

    def step(self, z, ctx):
        B = z.shape[0]
        zo, ls, ld = np.empty_like(z), np.empty_like(z), np.empty((B,), np.float32)
        _check(self.lib.iaf_step_fwd(self.plan, _p(z), _p(ctx), _p(zo), _p(ls), _p(ld), B, None))
        return zo, ls, ld

    def multiconv(self, z, ctx):
        B = z.shape[0]
        outs = [np.empty((B, h, self.H, self.W), np.float32) for h in self.heads]
        _check(self.lib.iaf_multiconv_fwd(self.plan, _p(z), _p(ctx), _arr(outs), B, None))
        return outs

    def layer(self, eps, post_mean, post_logsd, prior_mean, prior_logsd, ctx):
        B = eps.shape[0]
        f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        eps, post_mean, post_logsd, prior_mean, prior_logsd, ctx = map(f, (eps, post_mean, post_logsd, prior_mean, prior_logsd, ctx))
        zo, kl = np.empty_like(eps), np.empty_like(eps)
        kl_bc, kl_cost = np.empty((B, self.n_z), np.float32), np.empty((B,), np.float32)
        _check(self.lib.iaf_layer_fwd(self.plan, _p(eps), _p(post_mean), _p(post_logsd), _p(prior_mean), _p(prior_logsd),
                                      _p(ctx), _p(zo), _p(kl), _p(kl_bc), _p(kl_cost), B, None))
        return zo, kl, kl_bc, kl_cost

    def layer_bwd(self, eps, post_mean, post_logsd, prior_mean, prior_logsd, ctx, g_zout, g_kl, g_kl_bc, g_kl_cost,
                  params=True):
        B = eps.shape[0]
        _, g_ctx, gw, gs, gb = self._grad_bufs(eps, ctx, params)
        outs = [np.full_like(eps, np.nan) for _ in range(5)]  # g_post_mean, g_post_logsd, g_prior_mean, g_prior_logsd, g_eps
        _check(self.lib.iaf_layer_bwd(self.plan, _p(eps), _p(post_mean), _p(post_logsd), _p(prior_mean), _p(prior_logsd),
                                      _p(ctx), _arr([l[0] for l in self.layers]), _arr([l[1] for l in self.layers]),
                                      _p(g_zout), _p(g_kl), _p(g_kl_bc), _p(g_kl_cost), _p(outs[0]), _p(outs[1]),
                                      _p(outs[2]), _p(outs[3]), _p(outs[4]), _p(g_ctx), _arr(gw) if params else None,
                                      _arr(gs) if params else None, _arr(gb) if params else None, B, None))
        return outs, g_ctx, gw, gs, gb

    def _grad_bufs(self, z, ctx, params):
        g_z = np.full_like(z, np.nan)
        g_ctx = np.full_like(ctx, np.nan) if ctx is not None and self.hidden else None
        gw = gs = gb = None
        if params:
            gw = [np.full_like(l[0], np.nan) for l in self.layers]
            gs = [np.full_like(l[1], np.nan) for l in self.layers]
            gb = [np.full_like(l[2], np.nan) for l in self.layers]
        return g_z, g_ctx, gw, gs, gb

    def step_bwd(self, z, ctx, g_zout, g_logsd=None, g_logdet=None, params=True):
        B = z.shape[0]
        g_z, g_ctx, gw, gs, gb = self._grad_bufs(z, ctx, params)
        _check(self.lib.iaf_step_bwd(self.plan, _p(z), _p(ctx), _arr([l[0] for l in self.layers]),
                                     _arr([l[1] for l in self.layers]), _p(g_zout), _p(g_logsd), _p(g_logdet), _p(g_z),
                                     _p(g_ctx), _arr(gw) if params else None, _arr(gs) if params else None,
                                     _arr(gb) if params else None, B, None))
        return g_z, g_ctx, gw, gs, gb

    def step_train(self, z, ctx):
        B = z.shape[0]
        zo, ls, ld = np.empty_like(z), np.empty_like(z), np.empty((B,), np.float32)
        hidden = [np.full((B, h, self.H, self.W), np.nan, np.float32) for h in self.hidden]
        harr = _arr(hidden) if hidden else None
        _check(self.lib.iaf_step_fwd_train(self.plan, _p(z), _p(ctx), _p(zo), _p(ls), _p(ld), harr, B, None))
        return zo, ls, ld, hidden

    def step_bwd_saved(self, z, ctx_like, zo, ls, hidden, g_zout, g_logsd=None, g_logdet=None, params=True):
        B = z.shape[0]
        g_z, g_ctx, gw, gs, gb = self._grad_bufs(z, ctx_like, params)
        harr = _arr(hidden) if hidden else None
        _check(self.lib.iaf_step_bwd_saved(self.plan, _p(z), _p(zo), _p(ls), harr, _arr([l[0] for l in self.layers]),
                                           _arr([l[1] for l in self.layers]), _p(g_zout), _p(g_logsd), _p(g_logdet),
                                           _p(g_z), _p(g_ctx), _arr(gw) if params else None, _arr(gs) if params else None,
                                           _arr(gb) if params else None, B, None))
        return g_z, g_ctx, gw, gs, gb

    def multiconv_train(self, z, ctx):
        B = z.shape[0]
        outs = [np.empty((B, h, self.H, self.W), np.float32) for h in self.heads]
        hidden = [np.full((B, h, self.H, self.W), np.nan, np.float32) for h in self.hidden]
        _check(self.lib.iaf_multiconv_fwd_train(self.plan, _p(z), _p(ctx), _arr(outs), _arr(hidden) if hidden else None, B, None))
        return outs, hidden

    def multiconv_bwd_saved(self, z, ctx_like, hidden, g_outs, params=True):
        B = z.shape[0]
        g_z, g_ctx, gw, gs, gb = self._grad_bufs(z, ctx_like, params)
        _check(self.lib.iaf_multiconv_bwd_saved(self.plan, _p(z), _arr(hidden) if hidden else None,
                                                _arr([l[0] for l in self.layers]), _arr([l[1] for l in self.layers]),
                                                _arr(g_outs), _p(g_z), _p(g_ctx), _arr(gw) if params else None,
                                                _arr(gs) if params else None, _arr(gb) if params else None, B, None))
        return g_z, g_ctx, gw, gs, gb

    def multiconv_bwd(self, z, ctx, g_outs, params=True):
        B = z.shape[0]
        g_z, g_ctx, gw, gs, gb = self._grad_bufs(z, ctx, params)
        _check(self.lib.iaf_multiconv_bwd(self.plan, _p(z), _p(ctx), _arr([l[0] for l in self.layers]),
                                          _arr([l[1] for l in self.layers]), _arr(g_outs), _p(g_z), _p(g_ctx),
                                          _arr(gw) if params else None, _arr(gs) if params else None,
                                          _arr(gb) if params else None, B, None))
        return g_z, g_ctx, gw, gs, gb
