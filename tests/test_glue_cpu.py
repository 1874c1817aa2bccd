"""Host logic of the SURVEY 8(f) rows on the torch emulation of the kernel table (CPU): the label wire format through
Generator / Discriminator (prologue table + symmetrisation), the gradient-penalty glue, the fused readout + argmax and the
flat-bucket AdamW against torch.optim.AdamW.  The CUDA kernels themselves are covered by tests/test_glue_gpu.py."""
import pytest
import torch

import druggen_b200 as dg
from druggen_b200 import gan, kernels, ops
from druggen_b200.optim import FlatAdamW
from conftest import rel_l2
from emul_kernels import EmulBackend


@pytest.fixture(autouse=True)
def emul():
    kernels._install_backend_for_tests(EmulBackend())
    old = kernels.get_precision()
    kernels.set_precision("fp32")
    yield
    kernels.set_precision(old)
    kernels._install_backend_for_tests(None)


def _nets(n=5, depth=2, dim=128, seed=0, dtype=torch.float64):
    torch.manual_seed(seed)
    G = dg.Generator("relu", n, 5, 13, 0.0, dim=dim, depth=depth, heads=8, mlp_ratio=3).to(dtype)
    D = dg.Discriminator("relu", n, 5, 13, 0.0, dim=dim, depth=depth, heads=8, mlp_ratio=3).to(dtype)
    return G, D


@pytest.mark.parametrize("label_dtype", [torch.uint8, torch.int64])
def test_label_inputs_equal_onehot_inputs(label_dtype):
    """G / D called with integer labels [B,N,N] / [B,N] (the 1-byte wire format, SURVEY 8f rows 1+3) give the outputs AND
    every parameter gradient of the same call on the fp32 one-hot tensors load_molecules produces (models.py:91-94)."""
    G, D = _nets()
    a_l, x_l = gan.synthetic_molecules(3, 5, 13, 5, seed=4, labels=True)
    a, x = gan.synthetic_molecules(3, 5, 13, 5, seed=4)
    a, x = a.double(), x.double()
    outs = {}
    for kind, (ze, zn) in {"dense": (a, x), "labels": (a_l.to(label_dtype), x_l.to(label_dtype))}.items():
        G.zero_grad(set_to_none=True); D.zero_grad(set_to_none=True)
        node, edge, ns, es = G(ze, zn)
        score = D(ze, zn)
        ((ns ** 2).sum() + (es ** 2).sum() + score.sum()).backward()
        outs[kind] = ([node, edge, ns, es, score], {k: v.grad.clone() for k, v in list(G.named_parameters()) + list(D.named_parameters())
                                                    if v.grad is not None})
    for t0, t1 in zip(outs["dense"][0], outs["labels"][0]):
        assert rel_l2(t1, t0) < 1e-12
    assert outs["dense"][1].keys() == outs["labels"][1].keys()
    for k, g0 in outs["dense"][1].items():
        assert rel_l2(outs["labels"][1][k], g0) < 1e-10, k


def test_asymmetric_labels_are_symmetrised_like_the_dense_path():
    G, _ = _nets(depth=1)
    a_l = torch.randint(0, 5, (2, 5, 5), generator=torch.Generator().manual_seed(1)).to(torch.uint8)     # NOT symmetric
    x_l = torch.randint(0, 13, (2, 5), generator=torch.Generator().manual_seed(2)).to(torch.uint8)
    a = torch.nn.functional.one_hot(a_l.long(), 5).double()
    x = torch.nn.functional.one_hot(x_l.long(), 13).double()
    with torch.no_grad():
        for t0, t1 in zip(G(a, x), G(a_l, x_l)):
            assert rel_l2(t1, t0) < 1e-12


def test_decode_matches_readout_then_max():
    G, _ = _nets(depth=1, dtype=torch.float32)
    a_l, x_l = gan.synthetic_molecules(3, 5, 13, 5, seed=9, labels=True)
    with torch.no_grad():
        _, _, ns, es = G(a_l, x_l)
        n_idx, e_idx = G.decode(a_l, x_l)
        n8, e8 = G.decode(a_l, x_l, idx_dtype=torch.uint8)
    assert torch.equal(n_idx, torch.max(ns, -1)[1]) and torch.equal(e_idx, torch.max(es, -1)[1])     # inference.py:197-198
    assert n8.dtype == torch.uint8 and torch.equal(n8.long(), n_idx) and torch.equal(e8.long(), e_idx)


def test_gp_interp_and_penalty_equal_loss_py():
    """loss.py:21-26 and :42-47 restated with torch ops vs the glue primitives (values, and gradients of the penalty)."""
    g = torch.Generator().manual_seed(3)
    b, n = 4, 5
    a_l, x_l = gan.synthetic_molecules(b, n, 13, 5, seed=11, labels=True)
    a, x = gan.synthetic_molecules(b, n, 13, 5, seed=11)
    fake_e, fake_n = torch.randn(b, n, n, 5, generator=g), torch.randn(b, n, 13, generator=g)
    eps_e, eps_n = torch.rand(b, 1, 1, 1, generator=g), torch.rand(b, 1, 1, generator=g)
    assert torch.equal(kernels.gp_interp(a_l, fake_e, eps_e), eps_e * a + (1 - eps_e) * fake_e)
    assert torch.equal(kernels.gp_interp(x_l, fake_n, eps_n), eps_n * x + (1 - eps_n) * fake_n)
    gn = torch.randn(b, n, 13, generator=g, dtype=torch.float64, requires_grad=True)
    ge = torch.randn(b, n, n, 5, generator=g, dtype=torch.float64, requires_grad=True)
    want = ((torch.cat([gn.reshape(b, -1), ge.reshape(b, -1)], 1).norm(2, dim=1) - 1) ** 2).mean()
    wn, we = torch.autograd.grad(want * 3.0, [gn, ge])
    got = ops.GradPenalty.apply(gn, ge)
    hn, he = torch.autograd.grad(got * 3.0, [gn, ge])
    assert abs(got.item() - want.item()) < 1e-12 and rel_l2(hn, wn) < 1e-12 and rel_l2(he, we) < 1e-12


def test_trainer_step_with_labels_equals_step_with_onehots():
    """GANTrainer.step fed the label wire format takes the same D / G losses and lands on the same weights as fed the fp32
    one-hot tensors (same eps draws)."""
    res = []
    for labels in (False, True):
        G, D = _nets(depth=2, dim=128, dtype=torch.float32)
        tr = gan.GANTrainer(G, D, lr_g=1e-3, lr_d=1e-3)
        mol = gan.synthetic_molecules(4, 5, 13, 5, seed=7, labels=labels)
        drug = gan.synthetic_molecules(4, 5, 13, 5, seed=8, labels=labels)
        torch.manual_seed(5)
        losses = [tr.step(drug[0], drug[1], mol[0], mol[1]) for _ in range(2)]
        res.append((losses, [p.detach().clone() for p in list(G.parameters()) + list(D.parameters())]))
    for l0, l1 in zip(res[0][0], res[1][0]):
        assert abs(l0[0] - l1[0]) < 1e-4 * max(1.0, abs(l0[0])) and abs(l0[1] - l1[1]) < 1e-4 * max(1.0, abs(l0[1]))
    for p0, p1 in zip(res[0][1], res[1][1]):
        assert rel_l2(p1, p0) < 1e-4


def test_flat_adamw_equals_torch_adamw_incl_none_grads():
    """optim.FlatAdamW (one fused launch over flat buckets) vs torch.optim.AdamW(lr, betas) (train.py:213-214) over 5 steps,
    one tensor never receiving a gradient (stays untouched, no decay) and one receiving it only from step 3 on (its own
    bias-correction step count)."""
    torch.manual_seed(0)
    shapes = [(7, 5), (5,), (3, 4, 2), (6,), (9, 2)]
    ref = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
    ours = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    o_ref = torch.optim.AdamW(ref, 1e-2, (0.9, 0.999))
    o_our = FlatAdamW(ours, 1e-2, (0.9, 0.999))
    assert all(p.data_ptr() >= o_our.flat_p.data_ptr() for p in ours)      # parameters are views of the flat buffer
    for step in range(5):
        o_ref.zero_grad(set_to_none=True); o_our.zero_grad(set_to_none=True)
        for i, (a, b) in enumerate(zip(ref, ours)):
            if i == 3 or (i == 1 and step < 2):
                continue
            gr = torch.randn(a.shape)
            a.grad, b.grad = gr.clone(), gr.clone()
        o_ref.step(); o_our.step()
        for a, b in zip(ref, ours):
            assert rel_l2(b, a) < 1e-6, step
    assert torch.equal(ours[3].detach(), ref[3].detach())
