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

    def _embed(self, z_e, z_n):
        node = self.node_layers(z_n)                       # models.py:91
        edge = self.edge_layers(z_e)                       # models.py:92
        edge = (edge + edge.permute(0, 2, 1, 3)) / 2       # models.py:94
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
        return node, edge, self.readout_n(node), self.readout_e(edge)


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
        return self.node_mlp(node.reshape(z_n.shape[0], -1))
