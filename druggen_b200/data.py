"""Producer of the hot path's inputs: ``load_molecules`` (reference src/data/utils.py:128-143) on the device.

The reference densifies a PyG mini-batch with ``torch_geometric.utils.to_dense_adj`` and expands the integer bond labels to fp32
one-hots on the host side of every step.  Here the edge list goes to the device as it is (int64 ``edge_index`` / ``batch`` /
``edge_attr``), ``dg_to_dense_adj`` scatters it into [B, N, N] labels, and the model takes either the fp32 one-hots
(``load_molecules``: the reference's return values) or the 1-byte labels themselves (``load_molecule_labels``: the wire format
``Generator`` / ``Discriminator`` / ``GANTrainer`` accept, 20x fewer bytes than the one-hot).  ``data`` is duck-typed (``.x``,
``.edge_index``, ``.edge_attr``, ``.batch``): torch_geometric is not imported.
"""
from __future__ import annotations

import torch

from . import kernels as K


def _fields(data, device):
    x, ei, ea, batch = data.x, data.edge_index, data.edge_attr, data.batch
    if device is not None:
        x, ei, ea, batch = (t.to(device, non_blocking=True) for t in (x, ei, ea, batch))
    return x, ei, ea, batch


def load_molecule_labels(data=None, b_dim=32, m_dim=32, device=None, batch_size=32):
    """-> (bond labels uint8 [B, N, N], atom labels uint8 [B, N]): the label wire format of one mini-batch.
    N = nodes per graph = data.batch.shape[0] / batch_size (src/data/utils.py:134); atom labels are the argmax of the one-hot
    ``data.x`` rows (src/data/utils.py:136 reshapes them, the dataset stores them one-hot)."""
    x, ei, ea, batch = _fields(data, device)
    n = int(batch.shape[0] / batch_size)
    adj = K.to_dense_adj(ei, batch, ea.view(-1), max_num_nodes=n, batch_size=batch_size)
    bonds = K.narrow_labels(adj, b_dim, validate=False)
    atoms = K.argmax_last(x.view(batch_size, n, -1).float().contiguous()).to(torch.uint8)
    K.check_labels()
    return bonds, atoms


def load_molecules(data=None, b_dim=32, m_dim=32, device=None, batch_size=32):
    """src/data/utils.py:128-143, same signature and return values: (real_graphs [B, N m + N N b], a_tensor [B,N,N,b] fp32 one-hot,
    x_tensor [B,N,m])."""
    x, ei, ea, batch = _fields(data, device)
    n = int(batch.shape[0] / batch_size)
    adj = K.to_dense_adj(ei, batch, ea.view(-1), max_num_nodes=n, batch_size=batch_size)
    a_tensor = K.label2onehot(K.narrow_labels(adj, b_dim, validate=False), b_dim)            # (:137; raises on a label >= b_dim)
    x_tensor = x.view(batch_size, n, -1)                                                      # (:136)
    real_graphs = torch.concat((x_tensor.reshape(batch_size, -1), a_tensor.reshape(batch_size, -1)), dim=-1)   # (:139-141)
    return real_graphs, a_tensor, x_tensor
