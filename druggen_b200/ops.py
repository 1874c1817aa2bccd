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
        ctx.prec = K.get_precision()       # the derivatives of a contraction run in the precision mode its forward ran in
        return K.rows_gemm(a, w, w_is_nk)

    @staticmethod
    def backward(ctx, dout):
        a, w = ctx.saved_tensors
        dout = _c(dout)
        da = dw = None
        with K.precision(ctx.prec):
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
        ctx.prec = K.get_precision()
        return K.gemm_tn(a, b)

    @staticmethod
    def backward(ctx, dout):
        a, b = ctx.saved_tensors
        dout = _c(dout)
        da = db = None
        with K.precision(ctx.prec):
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
        ctx.prec = K.get_precision()
        ctx.save_for_backward(x, w, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = _c(dy)
        with K.precision(ctx.prec):
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


# ----------------------------------------------------------------------------- the two-layer MLP as one primitive
class MLP(Function):
    """m = fc2(relu(fc1(x) + b1)) + b2   (layers.py:51-53) as ONE twice-differentiable primitive.

    Same arithmetic as two ``Linear`` primitives; what it buys is storage: the hidden activation h and its gradient
    dh -- the widest tensors of the block ([rows, mlp_ratio*dim]) -- are only ever contraction operands or a sign
    mask, so with ``narrow`` (tensor-core mode) they live in HBM as bf16, and the ReLU mask is applied in the GEMM
    epilogue instead of a separate pass.  The gradient penalty's double backward runs on these primitives."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, narrow, residual):
        h = K.rows_gemm(x, w1, True, b1, relu=True, out_bf16=narrow)
        ctx.save_for_backward(x, h, w1, b1, w2)
        ctx.narrow, ctx.residual = narrow, residual
        return K.rows_gemm(h, w2, True, b2, resid=x if residual else None)      # residual: x + mlp(x), the add in the store

    @staticmethod
    def backward(ctx, dm):
        x, h, w1, b1, w2 = ctx.saved_tensors
        nx, nw1, nb1, nw2, nb2 = ctx.needs_input_grad[:5]
        dx, dw1, db1, dw2, db2 = MLPBwd.apply(_c(dm), x, h, w1, b1, w2, ctx.narrow, ctx.residual, nx, nw1 or nb1, nw2 or nb2)
        return dx, dw1, db1, dw2, db2, None, None


class MLPBwd(Function):
    """(dm, x, h, w1, b1, w2) -> (dx, dW1, db1, dW2, db2):  dh = (dm W2) * (h > 0);  dx = dh W1;  dW1 = dh^T x;  dW2 = dm^T h.
    Its backward is the hand-derived second order (the mask is piecewise constant):
      t = u_dx W1^T + x u_dW1^T + u_db1,  tM = t * (h > 0)
      d/d dm = tM W2^T + h u_dW2^T + u_db2      d/d W2 = dm^T tM       d/d W1 = dh^T u_dx + p^T x
      d/d x  = dh u_dW1 + p W1                  d/d b1 = colsum(p)      with p = (dm u_dW2) * (h > 0).
    ``residual`` (the primitive computes x + mlp(x)): dx = dm + dh W1 and d/d dm gains u_dx -- both adds ride in a GEMM store."""

    @staticmethod
    def forward(ctx, dm, x, h, w1, b1, w2, narrow, residual, want_x, want_p1, want_p2):
        ctx.set_materialize_grads(False)
        ctx.narrow, ctx.residual = narrow, residual
        if narrow and residual and want_x:          # both GEMMs, the mask and the residual add in ONE tcgen05 chain launch
            dx, dh = K.mlp_bwd_dgrad(dm, h, w1, w2)
        else:
            dh = K.rows_gemm(dm, w2, False, gate=h, out_bf16=narrow)
            dx = K.rows_gemm(dh, w1, False, resid=dm if residual else None) if want_x else None  # residual: dx = dm + dh W1
        ctx.save_for_backward(dm, x, h, dh, w1, w2)
        dw1 = db1 = dw2 = db2 = None
        if want_p1:
            dw1, db1 = torch.zeros_like(w1), torch.zeros_like(b1)
            K.gemm_tn(dh, x, out=dw1, colsum_a=db1)
        if want_p2:
            dw2, db2 = torch.zeros_like(w2), torch.zeros(w2.shape[0], dtype=w2.dtype, device=w2.device)
            K.gemm_tn(dm, h, out=dw2, colsum_a=db2)
        return dx, dw1, db1, dw2, db2

    @staticmethod
    @once_differentiable
    def backward(ctx, u_dx, u_dw1, u_db1, u_dw2, u_db2):
        dm, x, h, dh, w1, w2 = ctx.saved_tensors
        narrow = ctx.narrow
        g_dm = g_x = g_w1 = g_b1 = g_w2 = None
        u_dx, u_dw1, u_dw2 = _c(u_dx), _c(u_dw1), _c(u_dw2)
        tm = None
        if u_dx is not None and u_dw1 is None and u_db1 is None and narrow and ctx.residual:
            # the gradient-penalty case in the tensor-core mode: tM = (u W1^T) * M and d/d dm = u + tM W2^T are the dgrad chain
            # with the two weights transposed into each other's role -- one launch
            g_dm, tm = K.mlp_bwd_dgrad(u_dx, h, w2.t().contiguous(), w1.t().contiguous())
            g_w2 = K.gemm_tn(dm, tm)
            tm = None
        elif u_dx is not None and u_dw1 is None and u_db1 is None:        # the gradient-penalty case: one gated GEMM
            tm = K.rows_gemm(u_dx, w1, True, gate=h, out_bf16=narrow)
        elif u_dx is not None or u_dw1 is not None or u_db1 is not None:
            t = None
            for term in (K.rows_gemm(u_dx, w1, True) if u_dx is not None else None,
                         K.rows_gemm(x, u_dw1, True) if u_dw1 is not None else None, u_db1):
                if term is not None:
                    t = term if t is None else t + term
            if t.dim() == 1:
                t = t.unsqueeze(0).expand(dm.shape[0], -1)
            tm = K.gate_mul(t.contiguous(), h.to(t.dtype))
        res_u = u_dx if ctx.residual else None        # residual form: dx = dm + dh W1  =>  d/d dm += u_dx (added in the store)
        if tm is not None:
            g_dm = K.rows_gemm(tm, w2, True, resid=res_u)
            g_w2 = K.gemm_tn(dm, tm)
        elif res_u is not None and g_dm is None:
            g_dm = res_u
        if u_dx is not None:
            g_w1 = K.gemm_tn(dh, u_dx)
        if u_dw1 is not None:
            g_x = K.rows_gemm(dh, u_dw1, False)
        if u_dw2 is not None:
            hu = K.rows_gemm(h, u_dw2, True)
            g_dm = hu if g_dm is None else g_dm + hu
            p = K.rows_gemm(dm, u_dw2, False, gate=h, out_bf16=narrow)
            px = K.rows_gemm(p, w1, False)
            g_x = px if g_x is None else g_x + px
            pw = K.gemm_tn(p, x)
            g_w1 = pw if g_w1 is None else g_w1 + pw
            g_b1 = K.colsum(p.to(w1.dtype))
        if u_db2 is not None:
            ub = u_db2.unsqueeze(0).expand(dm.shape[0], -1)
            g_dm = ub.contiguous() if g_dm is None else g_dm + ub
        return g_dm, g_x, None, g_w1, g_b1, g_w2, None, None, None, None, None


def mlp(x, w1, b1, w2, b2, residual: bool = False):
    """fc2(relu(fc1(x))) (+ x when ``residual``) on the last dim; hidden stored as bf16 in the tensor-core mode."""
    shp = x.shape
    narrow = K.fused_available(shp[-1], w1.shape[0])
    y = MLP.apply(x.reshape(-1, shp[-1]), w1, b1, w2, b2, narrow, residual)
    return y.reshape(*shp[:-1], w2.shape[0])


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


# ----------------------------------------------------------------------------- either side of the encoder
class EmbedLabels(Function):
    """Prologue of a one-hot batch given as integer labels (models.py:91-94,196-199): ``lut`` [classes, D] is the prologue
    MLP applied to the identity (built with ordinary torch ops by the caller, so the weights get their gradients through it);
    forward = table rows (+ the symmetrisation of models.py:94 for edges), backward = a segmented row sum into ``dlut``."""

    @staticmethod
    def forward(ctx, lut, labels, sym):
        ctx.save_for_backward(labels)
        ctx.sym, ctx.classes = sym, lut.shape[0]
        return K.embed_labels_fwd(labels, lut.contiguous(), sym)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        (labels,) = ctx.saved_tensors
        return K.embed_labels_bwd(labels, _c(dy), ctx.classes, ctx.sym), None, None


class Symmetrize(Function):
    """models.py:94 for dense inputs: (e + e^T) / 2 over the two atom axes.  Linear and self-adjoint, so every order of its
    derivative is the same launch."""

    @staticmethod
    def forward(ctx, e):
        return K.symmetrize(e.contiguous())

    @staticmethod
    def backward(ctx, dy):
        return Symmetrize.apply(dy)


class GradPenalty(Function):
    """loss.py:42-47: (g_node [B,...], g_edge [B,...]) -> mean_b (|concat(g_node_b, g_edge_b)|_2 - 1)^2 in two small
    launches; its backward scales the two gradient tensors by 2 (|g_b| - 1) / (B |g_b|)."""

    @staticmethod
    def forward(ctx, g_node, g_edge):
        g_node, g_edge = g_node.contiguous(), g_edge.contiguous()
        pen, coef = K.gp_penalty(g_node, g_edge)
        ctx.save_for_backward(g_node, g_edge, coef)
        return pen.reshape(())

    @staticmethod
    @once_differentiable
    def backward(ctx, up):
        g_node, g_edge, coef = ctx.saved_tensors
        up = up.contiguous()
        return K.gp_penalty_bwd(g_node, coef, up), K.gp_penalty_bwd(g_edge, coef, up)
