"""-m gpu: the CUDA kernels of the data / metric rows (dg_to_dense_adj, dg_narrow_labels, dg_pack_bits, dg_tanimoto_agg), called
through the C-ABI, against the oracle (oracle/data_oracle.py), the reference-generated golden vectors, and -- at BASELINE's batch
size -- size-independent properties (symmetry, checksum of the edge attributes, self-similarity)."""
import types

import numpy as np
import pytest
import torch

from druggen_b200 import data as dgdata
from druggen_b200 import kernels as K
from druggen_b200 import metrics
from conftest import load_golden
from oracle import data_oracle as orc

pytestmark = pytest.mark.gpu


def _pyg(dev, *arrs):
    x, ei, ea, batch = (torch.from_numpy(a) for a in arrs)
    return types.SimpleNamespace(x=x, edge_index=ei, edge_attr=ea, batch=batch), dev


@pytest.mark.parametrize("B,N,seed", [(1, 1, 0), (6, 9, 3), (37, 45, 1), (300, 9, 2), (5, 90, 4)])
def test_to_dense_adj_bit_exact(cuda_dev, B, N, seed):
    x, ei, ea, batch = orc.synthetic_pyg_batch(B, N, seed=seed, p_bond=0.3)
    want = orc.to_dense_adj(ei, batch, ea, max_num_nodes=N, batch_size=B)
    t = lambda a: torch.from_numpy(a).to(cuda_dev)  # noqa: E731
    got = K.to_dense_adj(t(ei), t(batch), t(ea), max_num_nodes=N, batch_size=B)
    assert got.dtype == torch.int32 and np.array_equal(got.cpu().numpy(), want)
    # PyG's defaults (batch size and node count found on the device), no attributes (every edge counts one), duplicated edges add
    ei2 = np.concatenate([ei, ei[:, :5]], axis=1)
    got2 = K.to_dense_adj(t(ei2), t(batch))
    assert np.array_equal(got2.cpu().numpy(), orc.to_dense_adj(ei2, batch))
    # max_num_nodes below the graph size drops the edges that reach it (to_dense_adj's mask)
    if N > 2:
        got3 = K.to_dense_adj(t(ei), t(batch), t(ea), max_num_nodes=N - 2, batch_size=B)
        assert np.array_equal(got3.cpu().numpy(), orc.to_dense_adj(ei, batch, ea, max_num_nodes=N - 2, batch_size=B))
    # ragged graphs (PyG does not need equal sizes)
    keep = np.ones(batch.size, bool)
    keep[::7] = False
    if keep.sum() > 0 and B > 1:
        remap = np.cumsum(keep) - 1
        e_ok = keep[ei[0]] & keep[ei[1]]
        ei_r, ea_r, batch_r = remap[ei[:, e_ok]], ea[e_ok], batch[keep]
        got4 = K.to_dense_adj(t(ei_r), t(batch_r), t(ea_r), max_num_nodes=N, batch_size=B)
        assert np.array_equal(got4.cpu().numpy(), orc.to_dense_adj(ei_r, batch_r, ea_r, max_num_nodes=N, batch_size=B))
    assert np.array_equal(K.to_dense_adj(t(np.zeros((2, 0), np.int64)), t(batch), None, N, B).cpu().numpy(), np.zeros((B, N, N), np.int32))


def test_load_molecules_golden_and_error_behaviour(cuda_dev):
    g = load_golden("data_metric.npz")
    batch, dev = _pyg(cuda_dev, g["pyg_x"], g["pyg_edge_index"], g["pyg_edge_attr"], g["pyg_batch"])
    real, a_tensor, x_tensor = dgdata.load_molecules(batch, b_dim=5, m_dim=13, device=dev, batch_size=6)
    assert np.array_equal(a_tensor.cpu().numpy(), g["a_tensor"])                   # = the reference's label2onehot(to_dense_adj(..))
    assert torch.equal(x_tensor.cpu(), batch.x.view(6, 9, 13)) and real.shape == (6, 9 * 13 + 81 * 5)
    bonds, atoms = dgdata.load_molecule_labels(batch, b_dim=5, m_dim=13, device=dev, batch_size=6)
    assert bonds.dtype == torch.uint8 and np.array_equal(bonds.cpu().numpy(), g["adj_labels"])
    assert np.array_equal(atoms.cpu().numpy(), g["pyg_x"].argmax(1).reshape(6, 9))
    # a bond label >= b_dim: the reference's scatter_ raises inside label2onehot; so does this path
    with pytest.raises(RuntimeError):
        dgdata.load_molecules(batch, b_dim=3, m_dim=13, device=dev, batch_size=6)
    K.check_labels()                                                               # (flag cleared by the raise above)
    # a node id outside the batch
    bad = types.SimpleNamespace(x=batch.x, edge_index=batch.edge_index.clone(), edge_attr=batch.edge_attr, batch=batch.batch)
    bad.edge_index[0, 0] = 10 ** 6
    with pytest.raises(RuntimeError):
        dgdata.load_molecule_labels(bad, b_dim=5, m_dim=13, device=dev, batch_size=6)


def test_label_batch_feeds_the_model_like_the_one_hot(cuda_dev):
    """The wire format end to end: PyG batch -> dg_to_dense_adj -> 1-byte labels -> Discriminator == the reference's route
    (to_dense_adj -> label2onehot -> fp32 one-hots -> Discriminator)."""
    import druggen_b200 as dg
    x, ei, ea, batch = orc.synthetic_pyg_batch(4, 9, seed=5)
    pyg, dev = _pyg(cuda_dev, x, ei, ea, batch)
    torch.manual_seed(0)
    D = dg.Discriminator("relu", 9, 5, 13, 0.0, dim=128, depth=2, heads=8, mlp_ratio=3).to(dev)
    with dg.precision("fp32"), torch.no_grad():
        _, a_tensor, x_tensor = dgdata.load_molecules(pyg, b_dim=5, m_dim=13, device=dev, batch_size=4)
        bonds, atoms = dgdata.load_molecule_labels(pyg, b_dim=5, m_dim=13, device=dev, batch_size=4)
        want, got = D(a_tensor, x_tensor), D(bonds, atoms)
    assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max())


def test_full_batch_properties(cuda_dev):
    """BASELINE's batch (2048 molecules of 45 atoms): symmetric (both directions listed), zero diagonal, and the sum of the dense
    labels = the sum of the edge attributes (nothing dropped, nothing doubled)."""
    B, N = 2048, 45
    x, ei, ea, batch = orc.synthetic_pyg_batch(B, N, seed=9)
    t = lambda a: torch.from_numpy(a).to(cuda_dev)  # noqa: E731
    adj = K.to_dense_adj(t(ei), t(batch), t(ea), max_num_nodes=N, batch_size=B)
    assert torch.equal(adj, adj.transpose(1, 2)) and int(torch.diagonal(adj, dim1=1, dim2=2).abs().sum()) == 0
    assert int(adj.sum()) == int(ea.sum())
    lab = K.narrow_labels(adj, 5)
    assert torch.equal(lab.to(torch.int32), adj)


@pytest.mark.parametrize("F", [1024, 2048, 100, 64])
def test_tanimoto_vs_oracle(cuda_dev, F):
    rng = np.random.default_rng(F)
    stock = (rng.random((517, F)) < 0.08).astype(np.uint8)
    gen = (rng.random((260, F)) < 0.08).astype(np.uint8)
    stock[5] = 0; gen[9] = 0; gen[17] = stock[3]
    for agg in ("max", "mean"):
        for p in (1, 2):
            want = orc.average_agg_tanimoto(stock, gen, agg=agg, p=p, intdiv=True)
            got = metrics.average_agg_tanimoto(stock, gen, agg=agg, device=cuda_dev, p=p, intdiv=True)
            if agg == "max" and p == 1:
                assert np.array_equal(got, want), F                                 # bit-exact: same integers, one IEEE division
            else:
                assert np.allclose(got, want, rtol=2e-6, atol=0), (F, agg, p)
    assert metrics.average_agg_tanimoto(stock, gen, device=cuda_dev) == pytest.approx(orc.average_agg_tanimoto(stock, gen), rel=1e-7)
    # float inputs (the reference converts with .float()); few generated fingerprints against many stock ones (stock split over CTAs)
    got = metrics.average_agg_tanimoto(stock.astype(np.float32), gen[:3].astype(np.float32), device=cuda_dev, intdiv=True)
    assert np.array_equal(got, orc.average_agg_tanimoto(stock, gen[:3], intdiv=True))


def test_tanimoto_reference_golden_and_properties(cuda_dev):
    g = load_golden("data_metric.npz")
    stock, gen = np.unpackbits(g["fp_stock"], axis=1), np.unpackbits(g["fp_gen"], axis=1)
    assert np.array_equal(metrics.average_agg_tanimoto(stock, gen, agg="max", device=cuda_dev, intdiv=True), g["tan_max_p1"])
    for key, kw in (("tan_max_p2", dict(agg="max", p=2)), ("tan_mean_p1", dict(agg="mean")), ("tan_mean_p2", dict(agg="mean", p=2))):
        assert np.allclose(metrics.average_agg_tanimoto(stock, gen, device=cuda_dev, intdiv=True, **kw), g[key], rtol=2e-6, atol=0), key
    mean, std = metrics.internal_diversity(gen, device=cuda_dev)
    assert mean == pytest.approx(float(np.mean(1 - g["tan_self_mean"])), rel=1e-6)
    # size-independent property at a metric-sized set: every fingerprint's best match inside its own set is itself
    rng = np.random.default_rng(1)
    big = (rng.random((20000, 1024)) < 0.05).astype(np.uint8)
    assert np.all(metrics.average_agg_tanimoto(big, big, device=cuda_dev, intdiv=True) == 1.0)
