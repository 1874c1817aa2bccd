"""TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product).

CPU restatements, in numpy, of the two integer / byte-valued neighbours of the hot path:

* ``to_dense_adj`` -- torch_geometric.utils.to_dense_adj as ``load_molecules`` calls it (reference src/data/utils.py:130-135).
  torch_geometric is a THIRD-PARTY dependency of the reference (pinned ``torch-geometric==2.2.0``, environment.yml) that is not
  vendored under /root/reference and not installed here, so this part is a restatement of its published algorithm
  (torch_geometric/utils/to_dense_adj.py @2.2.0: per-graph node counts -> exclusive cumsum -> local indices -> mask by
  max_num_nodes -> scatter-add) and **its parity is unpinned**: it is anchored on the reference's call site and on hand-worked
  known-answer cases in tests/test_data_cpu.py, not on an execution of PyG.
* ``label2onehot`` / ``load_molecules`` -- src/data/utils.py:15-23,128-143.  Pinned: tests/golden/data_metric.npz holds the outputs
  of the reference's own ``label2onehot`` (its function text executed unmodified by oracle/make_golden_data.py).
* ``average_agg_tanimoto`` -- src/util/utils.py:566-611.  Pinned the same way (the module imports rdkit at the top and cannot be
  imported; the function's own text is executed unmodified).
"""
import numpy as np


def to_dense_adj(edge_index, batch, edge_attr=None, max_num_nodes=None, batch_size=None):
    edge_index, batch = np.asarray(edge_index, np.int64), np.asarray(batch, np.int64)
    if batch_size is None:
        batch_size = int(batch.max()) + 1 if batch.size else 1
    num_nodes = np.bincount(batch, minlength=batch_size)                       # scatter(one, batch, reduce='add')
    cum = np.concatenate([[0], np.cumsum(num_nodes)])
    idx0 = batch[edge_index[0]]
    idx1 = edge_index[0] - cum[batch][edge_index[0]]
    idx2 = edge_index[1] - cum[batch][edge_index[1]]
    if max_num_nodes is None:
        max_num_nodes = int(num_nodes.max())
    attr = np.ones(idx0.size, np.int64) if edge_attr is None else np.asarray(edge_attr, np.int64)
    keep = (idx1 < max_num_nodes) & (idx2 < max_num_nodes)
    idx0, idx1, idx2, attr = idx0[keep], idx1[keep], idx2[keep], attr[keep]
    adj = np.zeros(batch_size * max_num_nodes * max_num_nodes, np.int64)
    np.add.at(adj, idx0 * max_num_nodes * max_num_nodes + idx1 * max_num_nodes + idx2, attr)     # scatter(reduce='add')
    return adj.reshape(batch_size, max_num_nodes, max_num_nodes)


def label2onehot(labels, dim):
    """src/data/utils.py:15-23: zeros(list(labels.size()) + [dim]).scatter_(-1, labels.unsqueeze(-1), 1.)"""
    labels = np.asarray(labels, np.int64)
    if labels.size and (labels.min() < 0 or labels.max() >= dim):
        raise RuntimeError("index out of range in scatter_")
    out = np.zeros(labels.shape + (dim,), np.float32)
    np.put_along_axis(out, labels[..., None], 1.0, axis=-1)
    return out


def load_molecules(x, edge_index, edge_attr, batch, b_dim, batch_size):
    """src/data/utils.py:128-143 -> (real_graphs, a_tensor, x_tensor)."""
    n = int(batch.shape[0] / batch_size)
    a = to_dense_adj(edge_index, batch, edge_attr, max_num_nodes=n, batch_size=batch_size)
    x_tensor = np.asarray(x).reshape(batch_size, n, -1)
    a_tensor = label2onehot(a, b_dim)
    real = np.concatenate([x_tensor.reshape(batch_size, -1), a_tensor.reshape(batch_size, -1)], axis=-1)
    return real, a_tensor, x_tensor


def average_agg_tanimoto(stock_vecs, gen_vecs, batch_size=5000, agg="max", p=1, intdiv=False):
    """src/util/utils.py:566-611 with the torch.mm replaced by a numpy fp32 matmul (exact on 0/1 inputs below 2^24)."""
    assert agg in ["max", "mean"]
    agg_t = np.zeros(len(gen_vecs))
    total = np.zeros(len(gen_vecs))
    for j in range(0, stock_vecs.shape[0], batch_size):
        xs = np.asarray(stock_vecs[j:j + batch_size], np.float32)
        for i in range(0, gen_vecs.shape[0], batch_size):
            yg = np.asarray(gen_vecs[i:i + batch_size], np.float32).T
            tp = xs @ yg
            with np.errstate(invalid="ignore", divide="ignore"):
                jac = tp / (xs.sum(1, keepdims=True) + yg.sum(0, keepdims=True) - tp)
            jac[np.isnan(jac)] = 1
            if p != 1:
                jac = jac ** p
            if agg == "max":
                agg_t[i:i + yg.shape[1]] = np.maximum(agg_t[i:i + yg.shape[1]], jac.max(0))
            else:
                agg_t[i:i + yg.shape[1]] += jac.sum(0)
                total[i:i + yg.shape[1]] += jac.shape[0]
    if agg == "mean":
        agg_t /= total
    if p != 1:
        agg_t = agg_t ** (1 / p)
    return agg_t if intdiv else np.mean(agg_t)


def synthetic_pyg_batch(batch_size, n, m_dim=13, b_dim=5, seed=0, p_bond=None):
    """A mini-batch in PyG's collated layout, as DruggenDataset stores molecules (src/data/dataset.py): every graph padded to n
    nodes, one-hot atom rows x [B n, m_dim], both directions of every bond in edge_index [2, E] with bond labels edge_attr [E]
    (label 0 = no bond is never listed), batch [B n]."""
    rng = np.random.default_rng(seed)
    p_bond = (2.0 / n) if p_bond is None else p_bond
    src, dst, att = [], [], []
    for b in range(batch_size):
        iu, ju = np.triu_indices(n, 1)
        on = rng.random(iu.size) < p_bond
        lab = rng.integers(1, b_dim, on.sum())
        i, j = iu[on] + b * n, ju[on] + b * n
        src += [i, j]; dst += [j, i]; att += [lab, lab]
    edge_index = np.stack([np.concatenate(src), np.concatenate(dst)]).astype(np.int64)
    edge_attr = np.concatenate(att).astype(np.int64)
    perm = rng.permutation(edge_attr.size)                                    # (edge order is arbitrary)
    atoms = rng.integers(0, m_dim, batch_size * n)
    x = np.eye(m_dim, dtype=np.float32)[atoms]
    batch = np.repeat(np.arange(batch_size), n).astype(np.int64)
    return x, edge_index[:, perm], edge_attr[perm], batch
