"""Autograd wiring of druggen_b200/ops.py (first and second order) checked in fp64 on the torch
emulation of the kernel table -- validates the hand-derived second-order formulas that the
CUDA kernels implement.  CPU only."""
import pytest
import torch
from torch.autograd import gradcheck, gradgradcheck

from druggen_b200 import kernels, ops
from emul_kernels import EmulBackend


@pytest.fixture(autouse=True)
def emul():
    kernels._install_backend_for_tests(EmulBackend())
    old = kernels.get_precision()
    kernels.set_precision("fp32")          # the emulation computes exactly; bf16 storage would only add rounding
    yield
    kernels.set_precision(old)
    kernels._install_backend_for_tests(None)


def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return torch.randn(*shape, generator=g, dtype=torch.float64).requires_grad_(True)


def both(fn, inputs):
    assert gradcheck(fn, inputs, eps=1e-6, atol=1e-6)
    assert gradgradcheck(fn, inputs, eps=1e-6, atol=1e-6)


def test_linear():
    both(lambda x, w, b: ops.Linear.apply(x, w, b, False), (rnd(5, 6), rnd(4, 6), rnd(4)))


def test_linear_relu():
    x = rnd(5, 6)
    both(lambda x, w, b: ops.Linear.apply(x, w, b, True), (x, rnd(4, 6), rnd(4)))


def test_rows_gemm_and_tn():
    both(lambda a, w: ops.RowsGemm.apply(a, w, True), (rnd(5, 6), rnd(4, 6)))
    both(lambda a, w: ops.RowsGemm.apply(a, w, False), (rnd(5, 6), rnd(6, 4)))
    both(lambda a, b: ops.GemmTN.apply(a, b), (rnd(7, 3), rnd(7, 4)))


def test_add_ln():
    both(lambda a, b, g, be: ops.AddLN.apply(a, b, g, be), (rnd(4, 8), rnd(4, 8, seed=1), rnd(8), rnd(8, seed=2)))
    both(lambda a, g, be: ops.AddLN.apply(a, None, g, be), (rnd(4, 8), rnd(8), rnd(8, seed=2)))


def test_modulate():
    both(lambda q, k, e: ops.Modulate.apply(q, k, e, 0.25), (rnd(2, 3, 4), rnd(2, 3, 4, seed=1), rnd(2, 3, 3, 4)))


def test_softmax_agg():
    both(lambda a, v: ops.SoftmaxAgg.apply(a, v), (rnd(2, 3, 3, 4), rnd(2, 3, 4)))


def test_mlp_primitive():
    """the fused two-layer MLP primitive: first and second order incl. the weight / bias cotangent paths"""
    x, w1, b1, w2, b2 = rnd(5, 6), rnd(8, 6, seed=1), rnd(8, seed=2), rnd(6, 8, seed=3), rnd(6, seed=4)
    both(lambda x, w1, b1, w2, b2: ops.MLP.apply(x, w1, b1, w2, b2, False, False), (x, w1, b1, w2, b2))
    both(lambda x, w1, b1, w2, b2: ops.MLP.apply(x, w1, b1, w2, b2, False, True), (x, w1, b1, w2, b2))      # x + mlp(x)
    ref = torch.relu(x @ w1.t() + b1) @ w2.t() + b2
    assert torch.allclose(ops.mlp(x, w1, b1, w2, b2), ref, atol=1e-12)
    assert torch.allclose(ops.mlp(x, w1, b1, w2, b2, residual=True), ref + x, atol=1e-12)
