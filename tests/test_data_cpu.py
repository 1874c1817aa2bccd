"""Oracle of the data / metric rows (oracle/data_oracle.py) against the reference-generated golden vectors and hand-worked
known answers, and the host logic of druggen_b200/data.py + metrics.py on the torch emulation.  CPU only."""
import types

import numpy as np
import pytest
import torch

from druggen_b200 import data as dgdata
from druggen_b200 import kernels, metrics
from conftest import load_golden
from emul_kernels import EmulBackend
from oracle import data_oracle as orc


@pytest.fixture()
def emul():
    kernels._install_backend_for_tests(EmulBackend())
    yield
    kernels._install_backend_for_tests(None)


def test_to_dense_adj_known_answers():
    """torch_geometric.utils.to_dense_adj @2.2.0, worked by hand: two graphs of 3 and 2 nodes; a duplicated edge adds; an edge whose
    local index reaches max_num_nodes is dropped; without edge_attr every edge counts 1."""
    edge_index = np.array([[0, 1, 1, 2, 3, 4, 0, 2], [1, 0, 2, 1, 4, 3, 1, 0]])
    batch = np.array([0, 0, 0, 1, 1])
    attr = np.array([1, 1, 2, 2, 3, 3, 4, 1])
    a = orc.to_dense_adj(edge_index, batch, attr)
    want = np.zeros((2, 3, 3), np.int64)
    want[0, 0, 1] = 1 + 4; want[0, 1, 0] = 1; want[0, 1, 2] = 2; want[0, 2, 1] = 2; want[0, 2, 0] = 1
    want[1, 0, 1] = 3; want[1, 1, 0] = 3
    assert np.array_equal(a, want)
    a2 = orc.to_dense_adj(edge_index, batch, attr, max_num_nodes=2)          # node 2 of graph 0 falls outside
    assert a2.shape == (2, 2, 2) and np.array_equal(a2, want[:, :2, :2])
    ones = orc.to_dense_adj(edge_index, batch)
    assert ones.sum() == 8 and ones[0, 0, 1] == 2
    assert orc.to_dense_adj(np.zeros((2, 0), np.int64), batch).sum() == 0     # no edges at all


def test_oracle_label2onehot_and_load_molecules_vs_reference_golden():
    g = load_golden("data_metric.npz")
    assert np.array_equal(orc.label2onehot(g["adj_labels"], 5), g["a_tensor"])       # the reference's own label2onehot output
    real, a_tensor, x_tensor = orc.load_molecules(g["pyg_x"], g["pyg_edge_index"], g["pyg_edge_attr"], g["pyg_batch"], 5, 6)
    assert np.array_equal(a_tensor, g["a_tensor"]) and x_tensor.shape == (6, 9, 13) and real.shape == (6, 9 * 13 + 81 * 5)
    assert np.array_equal(a_tensor, a_tensor.transpose(0, 2, 1, 3))                   # both directions of every bond are listed
    with pytest.raises(RuntimeError):
        orc.label2onehot(np.array([[5]]), 5)                                          # scatter_ raises on an out-of-range label


def test_oracle_tanimoto_vs_reference_golden():
    g = load_golden("data_metric.npz")
    stock, gen = np.unpackbits(g["fp_stock"], axis=1), np.unpackbits(g["fp_gen"], axis=1)
    for agg in ("max", "mean"):
        for p in (1, 2):
            got = orc.average_agg_tanimoto(stock, gen, batch_size=128, agg=agg, p=p, intdiv=True)
            want = g[f"tan_{agg}_p{p}"]
            if agg == "max" and p == 1:
                assert np.array_equal(got, want)                                      # integers, one fp32 division, a max: exact
            else:
                assert np.allclose(got, want, rtol=2e-6, atol=0)
    assert orc.average_agg_tanimoto(stock, gen) == pytest.approx(float(g["tan_max_scalar"]), rel=1e-12)
    assert g["tan_max_p1"][11] == 1.0 and g["tan_max_p1"][3] == 1.0                   # the duplicated row; the empty row meets an empty row


def test_host_logic_on_the_emulation(emul):
    """druggen_b200.data.load_molecules / load_molecule_labels and metrics.average_agg_tanimoto wire the kernel table correctly
    (same return values as the reference functions); the CUDA kernels behind it are covered by tests/test_data_gpu.py."""
    g = load_golden("data_metric.npz")
    batch = types.SimpleNamespace(x=torch.from_numpy(g["pyg_x"]), edge_index=torch.from_numpy(g["pyg_edge_index"]),
                                  edge_attr=torch.from_numpy(g["pyg_edge_attr"]), batch=torch.from_numpy(g["pyg_batch"]))
    real, a_tensor, x_tensor = dgdata.load_molecules(batch, b_dim=5, m_dim=13, device=None, batch_size=6)
    assert np.array_equal(a_tensor.numpy(), g["a_tensor"]) and torch.equal(x_tensor, batch.x.view(6, 9, 13))
    assert real.shape == (6, 9 * 13 + 81 * 5) and torch.equal(real[:, :117], x_tensor.reshape(6, -1))
    bonds, atoms = dgdata.load_molecule_labels(batch, b_dim=5, m_dim=13, device=None, batch_size=6)
    assert bonds.dtype == torch.uint8 and np.array_equal(bonds.numpy(), g["adj_labels"])
    assert np.array_equal(atoms.numpy(), g["pyg_x"].argmax(1).reshape(6, 9))
    stock, gen = np.unpackbits(g["fp_stock"], axis=1), np.unpackbits(g["fp_gen"], axis=1)
    for agg in ("max", "mean"):
        for p in (1, 2):
            got = metrics.average_agg_tanimoto(stock, gen, agg=agg, device="cpu", p=p, intdiv=True)
            assert np.allclose(got, g[f"tan_{agg}_p{p}"], rtol=2e-6, atol=0), (agg, p)
    assert metrics.average_agg_tanimoto(stock, gen, device="cpu") == pytest.approx(float(g["tan_max_scalar"]), rel=1e-6)
    mean, std = metrics.internal_diversity(gen, device="cpu")
    assert mean == pytest.approx(float(np.mean(1 - g["tan_self_mean"])), rel=1e-6)


def test_data_path_refuses_cpu_without_the_extension():
    with pytest.raises(RuntimeError):
        kernels.to_dense_adj(torch.zeros(2, 0, dtype=torch.int64), torch.zeros(3, dtype=torch.int64), None, 3, 1)
    with pytest.raises(RuntimeError):
        kernels.pack_bits(torch.zeros(2, 64, dtype=torch.uint8))
