"""Host-side mirror of the reference encoder classes (src/model/layers.py) on the B200 kernels.

Same class names, constructor signatures, sub-module / parameter names (so reference
``state_dict`` checkpoints load unchanged: train.py:250-263, inference.py:135-139) and
``forward`` signatures; the arithmetic runs in the hand-written kernels of ``csrc/``.

  MLP                 reference layers.py:7-54
  MHA                 reference layers.py:56-137
  Encoder_Block       reference layers.py:139-193
  TransformerEncoder  reference layers.py:195-234
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops
from .block import BLOCK_PARAM_NAMES, encoder_block, encoder_forward_nograd


class MLP(nn.Module):
    """fc2(relu(fc1(x))); output dropout (layers.py:41-54)."""

    def __init__(self, in_feat, hid_feat=None, out_feat=None, dropout=0.):
        super().__init__()
        hid_feat = hid_feat or in_feat
        out_feat = out_feat or in_feat
        self.fc1 = nn.Linear(in_feat, hid_feat)
        self.act = nn.ReLU()
        self.fc2 = nn.Linear(hid_feat, out_feat)
        self.droprateout = nn.Dropout(dropout)

    def forward(self, x):
        h = ops.linear(x.contiguous(), self.fc1.weight, self.fc1.bias, relu=True)
        return self.droprateout(ops.linear(h, self.fc2.weight, self.fc2.bias))      # layers.py:54


class MHA(nn.Module):
    """Edge-modulated per-channel graph attention (layers.py:97-137)."""

    def __init__(self, dim, heads, attention_dropout=0.):
        super().__init__()
        assert dim % heads == 0
        self.heads = heads
        self.scale = 1. / math.sqrt(dim)   # kept for API parity; unused by the reference too (layers.py:84)
        self.q = nn.Linear(dim, dim)
        self.k = nn.Linear(dim, dim)
        self.v = nn.Linear(dim, dim)
        self.e = nn.Linear(dim, dim)
        self.d_k = dim // heads
        self.out_e = nn.Linear(dim, dim)
        self.out_n = nn.Linear(dim, dim)

    def forward(self, node, edge):
        node, edge = node.contiguous(), edge.contiguous()
        q = ops.linear(node, self.q.weight, self.q.bias)
        k = ops.linear(node, self.k.weight, self.k.bias)
        v = ops.linear(node, self.v.weight, self.v.bias)
        e = ops.linear(edge, self.e.weight, self.e.bias)
        a = ops.Modulate.apply(q, k, e, 1.0 / math.sqrt(self.d_k))
        edge_out = ops.linear(a, self.out_e.weight, self.out_e.bias)
        node_out = ops.linear(ops.SoftmaxAgg.apply(a, v), self.out_n.weight, self.out_n.bias)
        return node_out, edge_out


class Encoder_Block(nn.Module):
    """ln1 -> MHA -> residuals -> ln3/ln4 -> mlp/mlp2 residual -> ln5/ln6 (layers.py:174-193)."""

    def __init__(self, dim, heads, act, mlp_ratio=4, drop_rate=0.):
        super().__init__()
        self.ln1 = nn.LayerNorm(dim)
        self.attn = MHA(dim, heads, drop_rate)
        self.ln3 = nn.LayerNorm(dim)
        self.ln4 = nn.LayerNorm(dim)
        self.mlp = MLP(dim, dim * mlp_ratio, dim, dropout=drop_rate)
        self.mlp2 = MLP(dim, dim * mlp_ratio, dim, dropout=drop_rate)
        self.ln5 = nn.LayerNorm(dim)
        self.ln6 = nn.LayerNorm(dim)
        self._drop = drop_rate

    def _params(self):
        """The block's 30 tensors in BLOCK_PARAM_NAMES order, looked up by attribute on every call so it stays correct
        after .to()/.cuda(), load_state_dict and in nn.DataParallel replicas (whose parameters are plain attributes).  Written out
        (one attribute chain per sub-module instead of thirty dotted-name walks: 70 -> 30 us per call, which is half of a small
        graph-replayed encoder forward); ``tests/test_dropin_api.py`` holds it against the names."""
        a, m, m2 = self.attn, self.mlp, self.mlp2
        ln1, ln3, ln4, ln5, ln6 = self.ln1, self.ln3, self.ln4, self.ln5, self.ln6
        q, k, v, e, oe, on = a.q, a.k, a.v, a.e, a.out_e, a.out_n
        f1, f2, g1, g2 = m.fc1, m.fc2, m2.fc1, m2.fc2
        return [ln1.weight, ln1.bias, q.weight, q.bias, k.weight, k.bias, v.weight, v.bias, e.weight, e.bias, oe.weight, oe.bias,
                on.weight, on.bias, ln3.weight, ln3.bias, ln4.weight, ln4.bias, f1.weight, f1.bias, f2.weight, f2.bias,
                g1.weight, g1.bias, g2.weight, g2.bias, ln5.weight, ln5.bias, ln6.weight, ln6.bias]

    def forward(self, x, y, _edge_out: bool = True):
        # training-mode dropout (reference --dropout / --ddropout, default 0): layers.py:54 on both MLP outputs
        drop = self._drop if (self.training and self._drop > 0.0) else 0.0
        return encoder_block(x.contiguous(), y.contiguous(), self._params(), self.attn.heads, _edge_out, drop)


class TransformerEncoder(nn.Module):
    """``depth`` encoder blocks in sequence (layers.py:221-234)."""

    def __init__(self, dim, depth, heads, act, mlp_ratio=4, drop_rate=0.1):
        super().__init__()
        self.Encoder_Blocks = nn.ModuleList([
            Encoder_Block(dim, heads, act, mlp_ratio, drop_rate) for _ in range(depth)])
        # set by Discriminator: its last block's edge output has no consumer (models.py:202-207)
        self._discard_final_edge = False

    def forward(self, x, y):
        last = len(self.Encoder_Blocks) - 1
        blocks = list(self.Encoder_Blocks)
        no_graph = not torch.is_grad_enabled() or not (
            x.requires_grad or y.requires_grad or any(p.requires_grad for p in self.parameters()))
        if no_graph and blocks and not any(b.training and b._drop > 0.0 for b in blocks):
            # inference / no-grad passes: the whole encoder as one library call where the block-level entry points apply
            with torch.no_grad():
                return encoder_forward_nograd(x.contiguous(), y.contiguous(), [b._params() for b in blocks], blocks[0].attn.heads,
                                              not self._discard_final_edge)
        for i, block in enumerate(self.Encoder_Blocks):
            x, y = block(x, y, not (self._discard_final_edge and i == last))
        return x, y
