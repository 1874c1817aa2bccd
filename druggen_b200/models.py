"""Host-side mirror of the reference Generator / Discriminator (src/model/models.py:5-209).

Constructor signature ``(act, vertexes, edges, nodes, dropout, dim, depth, heads, mlp_ratio)``,
attributes, sub-module names and ``forward(z_e, z_n)`` return values are the reference's, so
``loss.py`` / ``train.py`` / ``inference.py`` run on these classes unchanged.  The encoder stack
runs on the sm_100a kernels; the prologue / readout / head Linears are small (K = 5, 13, 64) and
are SURVEY section 8(f) "next" rows.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import kernels as K
from . import ops
from .layers import TransformerEncoder


def _activation(act):
    table = {"relu": nn.ReLU, "leaky": nn.LeakyReLU, "sigmoid": nn.Sigmoid, "tanh": nn.Tanh}
    return table[act]() if isinstance(act, str) and act in table else act


class _GraphNet(nn.Module):
    def __init__(self, act, vertexes, edges, nodes, dropout, dim, depth, heads, mlp_ratio):
        super().__init__()
        self.vertexes, self.edges, self.nodes = vertexes, edges, nodes
        self.depth, self.dim, self.heads, self.mlp_ratio, self.dropout = depth, dim, heads, mlp_ratio, dropout
        act = _activation(act)
        self._act = act
        self.features = vertexes * vertexes * edges + vertexes * nodes
        self.transformer_dim = vertexes * vertexes * dim + vertexes * dim
        self.node_layers = nn.Sequential(nn.Linear(nodes, 64), act, nn.Linear(64, dim), act,
                                         nn.Dropout(self.dropout))
        self.edge_layers = nn.Sequential(nn.Linear(edges, 64), act, nn.Linear(64, dim), act,
                                         nn.Dropout(self.dropout))
        self.TransformerEncoder = TransformerEncoder(dim=dim, depth=depth, heads=heads, act=act,
                                                     mlp_ratio=mlp_ratio, drop_rate=dropout)

    def _on_kernels(self, x) -> bool:
        return (x.is_cuda or K._test_backend is not None) and x.dtype == torch.float32 and not (self.training and self.dropout > 0)

    def _linear(self, x, layer, act: bool = False):
        """``layer(x)`` (then the activation) on the twice-differentiable row-GEMM primitive (ops.linear): the small Linears
        either side of the encoder (K, N = 5 / 13 / 16 / 1 ...) are padded to multiples of 4 with zero rows / columns -- bit-for-bit
        neutral -- and a ReLU rides in the GEMM store.  No cuBLAS / cutlass launch is left on the path.  (The primitives remember
        the precision mode of their forward: every derivative of these layers, the gradient penalty's included, is fp32 too.)"""
        w, b = layer.weight, layer.bias
        n, k = w.shape
        pk, pn = (-k) % 4, (-n) % 4
        if pk:
            x, w = nn.functional.pad(x, (0, pk)), nn.functional.pad(w, (0, pk))
        if pn:
            w, b = nn.functional.pad(w, (0, 0, 0, pn)), nn.functional.pad(b, (0, pn))
        relu = act and isinstance(self._act, nn.ReLU)
        with K.precision("fp32"):          # a few MFLOP per molecule: fp32 FMA in every precision mode (as the reference's sgemm)
            y = ops.linear(x.contiguous(), w, b, relu=relu)
        if pn:
            y = y[..., :n]
        return self._act(y) if act and not relu else y

    def _seq(self, x, seq):
        """An ``nn.Sequential`` of Linear / activation / Dropout(p = 0 or eval) layers through ``_linear``."""
        mods = list(seq)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Linear):
                act = i + 1 < len(mods) and mods[i + 1] is self._act
                x = self._linear(x, m, act)
                i += 2 if act else 1
            else:
                x = m(x)
                i += 1
        return x

    def _embed(self, z_e, z_n):
        """Prologue + encoder.  ``z_e`` / ``z_n`` are the reference's fp32 tensors [B,N,N,edges] / [B,N,nodes], or -- the
        label wire format -- integer (uint8 / int64) labels [B,N,N] / [B,N] of one-hot molecules:
        the prologue MLP of a one-hot row is a row of a [classes, dim] table, so edges are written once as
        (T[a_ij] + T[a_ji]) / 2 straight from the 1-byte labels (``dg_embed_labels_fwd``)."""
        if not torch.is_floating_point(z_e):
            if self.dim == 128 and not (self.training and self.dropout > 0):
                w = self.edge_layers[0].weight
                eye_e = torch.eye(self.edges, dtype=w.dtype, device=w.device)
                eye_n = torch.eye(self.nodes, dtype=w.dtype, device=w.device)
                on_k = self._on_kernels(eye_e)
                lut_e = self._seq(eye_e, self.edge_layers) if on_k else self.edge_layers(eye_e)    # models.py:92 on the identity
                lut_n = self._seq(eye_n, self.node_layers) if on_k else self.node_layers(eye_n)    # models.py:91
                edge = ops.EmbedLabels.apply(lut_e, z_e, True)                                     # + models.py:94
                node = ops.EmbedLabels.apply(lut_n, z_n, False)
                return self.TransformerEncoder(node, edge)
            z_e = K.label2onehot(z_e, self.edges, validate=False)
            z_n = K.label2onehot(z_n, self.nodes, validate=False)
        node = self._seq(z_n, self.node_layers) if self._on_kernels(z_n) else self.node_layers(z_n)     # models.py:91
        if (isinstance(self._act, nn.ReLU) and (z_e.is_cuda or K._test_backend is not None) and not (self.training and self.dropout > 0)
                and self.dim % 4 == 0):
            # dense inputs (generated / interpolated molecules): both prologue Linears with their ReLU in the GEMM store
            # (twice-differentiable primitives: the gradient penalty differentiates through here), 64 -> dim on tcgen05 in the
            # tensor-core modes, and models.py:94 as one pass instead of permute + add + div
            l1, l2 = self.edge_layers[0], self.edge_layers[2]
            pad = (-self.edges) % 4                          # (the row GEMMs take K % 4 == 0: b_dim = 5 -> three zero columns)
            zp, w1 = (nn.functional.pad(z_e, (0, pad)), nn.functional.pad(l1.weight, (0, pad))) if pad else (z_e, l1.weight)
            h = ops.linear(zp.contiguous(), w1, l1.bias, relu=True)
            edge = ops.Symmetrize.apply(ops.linear(h, l2.weight, l2.bias, relu=True))
        else:
            edge = self.edge_layers(z_e)                   # models.py:92
            edge = (edge + edge.permute(0, 2, 1, 3)) / 2   # models.py:94
        return self.TransformerEncoder(node, edge)


class Generator(_GraphNet):
    """models.py:5-103.  forward -> (node, edge, node_sample, edge_sample), raw logits."""

    def __init__(self, act, vertexes, edges, nodes, dropout, dim, depth, heads, mlp_ratio):
        super().__init__(act, vertexes, edges, nodes, dropout, dim, depth, heads, mlp_ratio)
        self.readout_e = nn.Linear(self.dim, edges)
        self.readout_n = nn.Linear(self.dim, nodes)
        self.softmax = nn.Softmax(dim=-1)                  # defined, never applied (models.py:69)

    def forward(self, z_e, z_n):
        node, edge = self._embed(z_e, z_n)
        if self._on_kernels(node):                         # models.py:100-101 on the row-GEMM primitive (N = 13 / 5 padded to 16 / 8)
            return node, edge, self._linear(node, self.readout_n), self._linear(edge, self.readout_e)
        return node, edge, self.readout_n(node), self.readout_e(edge)

    @torch.no_grad()
    def decode(self, z_e, z_n, idx_dtype=torch.int64):
        """inference.py:195-198 in one call: G forward, then the readouts fused with ``torch.max(.., -1)[1]``
        (``dg_readout_argmax``: the [B,N,N,dim] stream is read once, the logits never reach HBM) ->
        (node labels [B,N], edge labels [B,N,N]) as int64, or uint8 for a 1-byte-per-edge result."""
        node, edge = self._embed(z_e, z_n)
        if self.dim != 128:
            return torch.max(self.readout_n(node), -1)[1], torch.max(self.readout_e(edge), -1)[1]
        n_idx, _ = K.readout_argmax(node.contiguous(), self.readout_n.weight, self.readout_n.bias, idx_dtype=idx_dtype)
        e_idx, _ = K.readout_argmax(edge.contiguous(), self.readout_e.weight, self.readout_e.bias, idx_dtype=idx_dtype)
        return n_idx, e_idx


class Discriminator(_GraphNet):
    """models.py:106-209.  forward -> [B,1] critic score from the flattened node stream."""

    def __init__(self, act, vertexes, edges, nodes, dropout, dim, depth, heads, mlp_ratio):
        super().__init__(act, vertexes, edges, nodes, dropout, dim, depth, heads, mlp_ratio)
        self.node_features = vertexes * dim
        self.edge_features = vertexes * vertexes * dim
        act = self._act
        self.node_mlp = nn.Sequential(nn.Linear(self.node_features, 64), act, nn.Linear(64, 32), act,
                                      nn.Linear(32, 16), act, nn.Linear(16, 1))
        self.TransformerEncoder._discard_final_edge = True

    def forward(self, z_e, z_n):
        node, _ = self._embed(z_e, z_n)
        flat = node.reshape(z_n.shape[0], -1)
        return self._seq(flat, self.node_mlp) if self._on_kernels(flat) else self.node_mlp(flat)          # models.py:207-208
