"""Differentiable primitives of the encoder block, closed under differentiation.

Each primitive is a ``torch.autograd.Function`` whose forward is one raw kernel from
``kernels.py`` and whose backward is expressed with primitives of the same set, so
``autograd.grad(..., create_graph=True)`` -- the WGAN-GP gradient penalty in the
reference's loss.py:32-39 -- differentiates straight through them.  The innermost
(second-order) backward of each ``*Bwd`` function is a hand-derived kernel
(formulas: DESIGN.md section "second-order formulas"; checked by gradgradcheck in
tests/test_ops_autograd.py).  Third-order derivatives are not provided.
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import kernels as K


def _c(t):
    return None if t is None else t.contiguous()


# ----------------------------------------------------------------------------- dense contractions
class RowsGemm(Function):
    """out = a . op(w)   (a:[R,K]; w:[N,K] if w_is_nk else [K,N])."""

    @staticmethod
    def forward(ctx, a, w, w_is_nk):
        ctx.save_for_backward(a, w)
        ctx.w_is_nk = w_is_nk
        return K.rows_gemm(a, w, w_is_nk)

    @staticmethod
    def backward(ctx, dout):
        a, w = ctx.saved_tensors
        dout = _c(dout)
        da = dw = None
        if ctx.needs_input_grad[0]:
            da = RowsGemm.apply(dout, w, not ctx.w_is_nk)
        if ctx.needs_input_grad[1]:
            dw = GemmTN.apply(dout, a) if ctx.w_is_nk else GemmTN.apply(a, dout)
        return da, dw, None


class GemmTN(Function):
    """out[M,N] = a[R,M]^T . b[R,N]."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return K.gemm_tn(a, b)

    @staticmethod
    def backward(ctx, dout):
        a, b = ctx.saved_tensors
        dout = _c(dout)
        da = db = None
        if ctx.needs_input_grad[0]:
            da = RowsGemm.apply(b, dout, True)
        if ctx.needs_input_grad[1]:
            db = RowsGemm.apply(a, dout, False)
        return da, db


class ColSum(Function):
    @staticmethod
    def forward(ctx, a):
        ctx.rows = a.shape[0]
        return K.colsum(a)

    @staticmethod
    def backward(ctx, dout):
        return dout.unsqueeze(0).expand(ctx.rows, -1).contiguous()


class GateMul(Function):
    """x * (ref > 0): the ReLU derivative applied to a gradient."""

    @staticmethod
    def forward(ctx, x, ref):
        ctx.save_for_backward(ref)
        return K.gate_mul(x, ref)

    @staticmethod
    def backward(ctx, dout):
        (ref,) = ctx.saved_tensors
        return GateMul.apply(_c(dout), ref), None


class Linear(Function):
    """y = act(x W^T + b) on rows, bias (+ReLU) fused in the GEMM epilogue (layers.py:51-53,111-135)."""

    @staticmethod
    def forward(ctx, x, w, b, relu):
        y = K.rows_gemm(x, w, True, b, relu)
        ctx.relu = relu
        ctx.save_for_backward(x, w, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = _c(dy)
        if ctx.relu:
            dy = GateMul.apply(dy, y)
        dx = RowsGemm.apply(dy, w, False) if ctx.needs_input_grad[0] else None
        dw = GemmTN.apply(dy, x) if ctx.needs_input_grad[1] else None
        db = ColSum.apply(dy) if ctx.needs_input_grad[2] else None
        return dx, dw, db, None


def linear(x, w, b, relu: bool = False):
    shp = x.shape
    y = Linear.apply(x.reshape(-1, shp[-1]), w, b, relu)
    return y.reshape(*shp[:-1], w.shape[0])


# ----------------------------------------------------------------------------- residual + LayerNorm
class AddLN(Function):
    """LN(a + b) (b optional) -- the four residual+LayerNorm sites of layers.py:187-192 and ln1."""

    @staticmethod
    def forward(ctx, a, b, gamma, beta):
        ctx.save_for_backward(a, b, gamma)
        return K.add_ln_fwd(a, b, gamma, beta)

    @staticmethod
    def backward(ctx, dy):
        a, b, gamma = ctx.saved_tensors
        dz, dgamma, dbeta = AddLNBwd.apply(_c(dy), a, b, gamma)
        return dz, (dz if b is not None else None), dgamma, dbeta


class AddLNBwd(Function):
    @staticmethod
    def forward(ctx, dy, a, b, gamma):
        ctx.save_for_backward(dy, a, b, gamma)
        ctx.set_materialize_grads(False)
        return K.add_ln_bwd(dy, a, b, gamma)

    @staticmethod
    @once_differentiable
    def backward(ctx, u, vg, vb):
        dy, a, b, gamma = ctx.saved_tensors
        if u is None:
            u = torch.zeros_like(a)
        g_dy, g_z, g_gamma = K.add_ln_bwd_bwd(_c(u), _c(vg), _c(vb), dy, a, b, gamma)
        return g_dy, g_z, (g_z if b is not None else None), g_gamma


def add_ln(a, b, gamma, beta):
    shp = a.shape
    d = shp[-1]
    y = AddLN.apply(a.reshape(-1, d), None if b is None else b.reshape(-1, d), gamma, beta)
    return y.reshape(shp)


# ----------------------------------------------------------------------------- edge-modulated scores
class Modulate(Function):
    @staticmethod
    def forward(ctx, q, k, e, c):
        ctx.save_for_backward(q, k, e)
        ctx.c = c
        return K.modulate_fwd(q, k, e, c)

    @staticmethod
    def backward(ctx, da):
        q, k, e = ctx.saved_tensors
        dq, dk, de = ModulateBwd.apply(_c(da), q, k, e, ctx.c)
        return dq, dk, de, None


class ModulateBwd(Function):
    @staticmethod
    def forward(ctx, da, q, k, e, c):
        ctx.save_for_backward(da, q, k, e)
        ctx.c = c
        return K.modulate_bwd(da, q, k, e, c)

    @staticmethod
    @once_differentiable
    def backward(ctx, uq, uk, ue):
        da, q, k, e = ctx.saved_tensors
        g_da, g_q, g_k, g_e = K.modulate_bwd_bwd(_c(uq), _c(uk), _c(ue), da, q, k, e, ctx.c)
        return g_da, g_q, g_k, g_e, None


# ----------------------------------------------------------------------------- softmax over keys + aggregate
class SoftmaxAgg(Function):
    @staticmethod
    def forward(ctx, a, v):
        ctx.save_for_backward(a, v)
        return K.softmax_agg_fwd(a, v)

    @staticmethod
    def backward(ctx, dg):
        a, v = ctx.saved_tensors
        return SoftmaxAggBwd.apply(_c(dg), a, v)


class SoftmaxAggBwd(Function):
    @staticmethod
    def forward(ctx, dg, a, v):
        ctx.save_for_backward(dg, a, v)
        return K.softmax_agg_bwd(dg, a, v)

    @staticmethod
    @once_differentiable
    def backward(ctx, ua, uv):
        dg, a, v = ctx.saved_tensors
        return K.softmax_agg_bwd_bwd(_c(ua), _c(uv), dg, a, v)
