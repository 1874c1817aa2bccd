"""Either side of the encoder path: label2onehot (src/data/utils.py:15-23) and the argmax decode (inference.py:197-198),
bit-exact against the reference statements run by ATen on the CPU."""
import pytest
import torch

import druggen_b200 as dg

pytestmark = pytest.mark.gpu


def ref_label2onehot(labels, dim):                 # src/data/utils.py:15-23, verbatim semantics
    out = torch.zeros(list(labels.size()) + [dim])
    out.scatter_(len(out.size()) - 1, labels.unsqueeze(-1), 1.)
    return out.float()


@pytest.mark.parametrize("shape,dim", [((2, 9, 9), 5), ((3, 45, 45), 5), ((7, 45), 13), ((1,), 5), ((0, 4), 5), ((64, 90, 90), 5)])
@pytest.mark.parametrize("dtype", [torch.int64, torch.uint8])
def test_label2onehot_bit_exact(cuda_dev, shape, dim, dtype):
    g = torch.Generator().manual_seed(sum(shape) + dim)
    labels = torch.randint(0, dim, shape, generator=g)
    want = ref_label2onehot(labels, dim)
    got = dg.label2onehot(labels.to(dtype).to(cuda_dev), dim)
    assert got.dtype == torch.float32 and got.shape == want.shape and torch.equal(got.cpu(), want)
    got2 = dg.label2onehot(labels.to(dtype), dim, device=cuda_dev)          # reference signature: host labels + device
    assert torch.equal(got2.cpu(), want)


def test_label2onehot_feeds_the_generator_like_the_dense_input(cuda_dev):
    """the 1-byte wire format expands to exactly the tensor synthetic_molecules / load_molecules hand to G"""
    from druggen_b200 import gan
    a, x = gan.synthetic_molecules(4, 9, 13, 5, seed=3)
    lab_a, lab_x = a.argmax(-1).to(torch.uint8), x.argmax(-1).to(torch.uint8)
    assert torch.equal(dg.label2onehot(lab_a, 5, device=cuda_dev).cpu(), a)
    assert torch.equal(dg.label2onehot(lab_x, 13, device=cuda_dev).cpu(), x)


@pytest.mark.parametrize("rows,C", [(1, 5), (2025, 5), (45, 13), (100003, 5), (0, 5)])
def test_argmax_decode_bit_exact_incl_ties_and_nan(cuda_dev, rows, C):
    g = torch.Generator().manual_seed(rows + C)
    t = torch.randn(rows, C, generator=g)
    tq = torch.round(t * 2) / 2                                               # quantised: many exact ties
    for x in (t, tq):
        assert torch.equal(dg.argmax_last(x.to(cuda_dev)).cpu(), torch.max(x, -1)[1])
    if rows > 3:
        tn = t.clone()
        tn[1, C - 1] = float("nan")
        tn[2, 0] = float("nan"); tn[2, 2] = float("nan")
        tn[3, :] = float("-inf")
        assert torch.equal(dg.argmax_last(tn.to(cuda_dev)).cpu(), torch.max(tn, -1)[1])
    shaped = t.view(-1, 1, C) if rows else t.view(0, 1, C)
    assert dg.argmax_last(shaped.to(cuda_dev)).shape == shaped.shape[:-1]


@pytest.mark.parametrize("dtype", [torch.int64, torch.uint8])
def test_label2onehot_rejects_out_of_range_like_scatter(cuda_dev, dtype):
    """src/data/utils.py:21 scatter_ raises on an index outside [0, dim); so does the kernel path (device flag -> RuntimeError)"""
    labels = torch.tensor([[0, 1, 7, 2]], dtype=dtype)
    with pytest.raises(RuntimeError):
        ref_label2onehot(labels.long(), 5)
    with pytest.raises(RuntimeError):
        dg.label2onehot(labels.to(cuda_dev), 5)
    dg.label2onehot(torch.tensor([[0, 4]], dtype=dtype, device=cuda_dev), 5)         # the flag was cleared by the raise
    deferred = dg.label2onehot(labels.to(cuda_dev), 5, validate=False)               # no sync here ...
    assert deferred.shape == (1, 4, 5)
    with pytest.raises(RuntimeError):
        dg.kernels.check_labels()                                                     # ... the error surfaces at the next check
