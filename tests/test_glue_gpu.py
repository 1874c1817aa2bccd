"""The CUDA kernels of the SURVEY 8(f) rows (csrc/glue.cu) through the C-ABI against plain-torch statements of the reference
lines they replace: bit-exact for the label / index work and the interpolation, 1e-5 for the fp32 reductions."""
import pytest
import torch

import druggen_b200 as dg
from druggen_b200 import gan, kernels as K, ops
from druggen_b200.optim import FlatAdamW
from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,N", [(1, 1), (3, 9), (5, 45), (2, 90), (0, 9)])
@pytest.mark.parametrize("dtype", [torch.uint8, torch.int64])
def test_embed_labels_fwd_bwd(cuda_dev, B, N, dtype):
    g = torch.Generator().manual_seed(B * 100 + N)
    for classes, sym in ((5, True), (13, False)):
        shape = (B, N, N) if sym else (B, N)
        lab = torch.randint(0, classes, shape, generator=g)            # deliberately NOT symmetric
        lut = torch.randn(classes, 128, generator=g)
        t = lut[lab]
        want = (t + t.transpose(1, 2)) / 2 if sym else t               # models.py:94 on table rows
        got = K.embed_labels_fwd(lab.to(dtype).to(cuda_dev), lut.to(cuda_dev), sym)
        assert got.shape == want.shape and torch.equal(got.cpu(), want)
        dy = torch.randn(*shape, 128, generator=g)
        lut64 = lut.double().requires_grad_(True)
        t64 = lut64[lab]
        (((t64 + t64.transpose(1, 2)) / 2 if sym else t64) * dy.double()).sum().backward() if B else None
        dl = K.embed_labels_bwd(lab.to(dtype).to(cuda_dev), dy.to(cuda_dev), classes, sym)
        if B:
            assert rel_l2(dl, lut64.grad) < 1e-5
        else:
            assert float(dl.abs().max()) == 0.0


def test_label_batch_through_generator_and_discriminator(cuda_dev):
    """G / D on uint8 labels == G / D on the fp32 one-hots (fp32 parity mode): outputs to 1e-5, parameter grads to 1e-4."""
    torch.manual_seed(0)
    G = dg.Generator("relu", 9, 5, 13, 0.0, dim=128, depth=2, heads=8, mlp_ratio=3).to(cuda_dev)
    D = dg.Discriminator("relu", 9, 5, 13, 0.0, dim=128, depth=2, heads=8, mlp_ratio=3).to(cuda_dev)
    a_l, x_l = gan.synthetic_molecules(6, 9, 13, 5, seed=4, labels=True, device=cuda_dev)
    a, x = gan.synthetic_molecules(6, 9, 13, 5, seed=4, device=cuda_dev)
    res = {}
    with dg.precision("fp32"):
        for kind, (ze, zn) in {"dense": (a, x), "labels": (a_l, x_l)}.items():
            G.zero_grad(set_to_none=True); D.zero_grad(set_to_none=True)
            node, edge, ns, es = G(ze, zn)
            score = D(ze, zn)
            ((ns ** 2).mean() + (es ** 2).mean() + score.mean()).backward()
            res[kind] = ([node, edge, ns, es, score], {k: v.grad.clone() for k, v in list(G.named_parameters()) + list(D.named_parameters())
                                                       if v.grad is not None})
    for t0, t1 in zip(res["dense"][0], res["labels"][0]):
        assert rel_l2(t1, t0) < 1e-5
    assert res["dense"][1].keys() == res["labels"][1].keys()
    for k, g0 in res["dense"][1].items():
        assert rel_l2(res["labels"][1][k], g0) < 1e-4, k
    K.check_labels()


@pytest.mark.parametrize("B,N", [(1, 9), (7, 45)])
def test_gp_interp_bit_exact_and_penalty(cuda_dev, B, N):
    g = torch.Generator().manual_seed(B + N)
    a_l, x_l = gan.synthetic_molecules(B, N, 13, 5, seed=11, labels=True, device=cuda_dev)
    a, x = gan.synthetic_molecules(B, N, 13, 5, seed=11, device=cuda_dev)
    fake_e, fake_n = torch.randn(B, N, N, 5, generator=g).to(cuda_dev), torch.randn(B, N, 13, generator=g).to(cuda_dev)
    eps_e, eps_n = torch.rand(B, 1, 1, 1, generator=g).to(cuda_dev), torch.rand(B, 1, 1, generator=g).to(cuda_dev)
    assert torch.equal(K.gp_interp(a_l, fake_e, eps_e), eps_e * a + (1 - eps_e) * fake_e)              # loss.py:25-26, bit for bit
    assert torch.equal(K.gp_interp(x_l, fake_n, eps_n), eps_n * x + (1 - eps_n) * fake_n)
    assert torch.equal(K.gp_interp(a_l.long(), fake_e, eps_e), eps_e * a + (1 - eps_e) * fake_e)
    gn = (torch.randn(B, N, 13, generator=g) * 0.05).to(cuda_dev).requires_grad_(True)
    ge = (torch.randn(B, N, N, 5, generator=g) * 0.01).to(cuda_dev).requires_grad_(True)
    want = ((torch.cat([gn.reshape(B, -1), ge.reshape(B, -1)], 1).double().norm(2, dim=1) - 1) ** 2).mean()   # loss.py:42-47
    wn, we = torch.autograd.grad(want * 3.0, [gn, ge])
    got = ops.GradPenalty.apply(gn, ge)
    hn, he = torch.autograd.grad(got * 3.0, [gn, ge])
    assert abs(got.item() - want.item()) < 1e-5 * max(1.0, abs(want.item()))
    assert rel_l2(hn, wn) < 1e-5 and rel_l2(he, we) < 1e-5


@pytest.mark.parametrize("rows,classes", [(1, 5), (2025, 5), (45, 13), (100003, 5), (0, 13)])
def test_readout_argmax(cuda_dev, rows, classes):
    g = torch.Generator().manual_seed(rows + classes)
    x, w, b = torch.randn(rows, 128, generator=g), torch.randn(classes, 128, generator=g) * 0.1, torch.randn(classes, generator=g) * 0.1
    want = (x.double() @ w.double().t() + b.double())
    idx, logits = K.readout_argmax(x.to(cuda_dev), w.to(cuda_dev), b.to(cuda_dev), want_logits=True)
    idx8, none = K.readout_argmax(x.to(cuda_dev), w.to(cuda_dev), b.to(cuda_dev), idx_dtype=torch.uint8)
    assert none is None and idx8.dtype == torch.uint8 and torch.equal(idx8.long(), idx)
    if rows == 0:
        return
    assert rel_l2(logits, want) < 1e-5
    assert torch.equal(idx.cpu(), torch.max(logits.cpu(), -1)[1])             # the index IS the first maximum of the logits it wrote
    top2 = want.topk(2, -1).values
    safe = (top2[:, 0] - top2[:, 1]) > 1e-4
    assert torch.equal(idx.cpu()[safe], want.argmax(-1)[safe])               # and the oracle's away from ties
    # exact ties: x = 0 -> logits = bias; equal biases -> the FIRST index
    idx0, _ = K.readout_argmax(torch.zeros(4, 128, device=cuda_dev), w.to(cuda_dev), torch.full((classes,), 0.5, device=cuda_dev))
    assert int(idx0.abs().max()) == 0


def test_flat_adamw_equals_torch_adamw(cuda_dev):
    torch.manual_seed(0)
    shapes = [(384, 128), (128,), (64, 5760), (3, 4, 2), (6,), (1,)]
    ref = [torch.nn.Parameter(torch.randn(s, device=cuda_dev)) for s in shapes]
    ours = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    o_ref = torch.optim.AdamW(ref, 1e-3, (0.9, 0.999), foreach=False)
    o_our = FlatAdamW(ours, 1e-3, (0.9, 0.999))
    for step in range(4):
        o_ref.zero_grad(set_to_none=True); o_our.zero_grad(set_to_none=True)
        for i, (a, b) in enumerate(zip(ref, ours)):
            if i == 4 or (i == 1 and step < 2):
                continue
            gr = torch.randn(a.shape, device=cuda_dev)
            a.grad, b.grad = gr.clone(), gr.clone()
        o_ref.step(); o_our.step()
        for a, b in zip(ref, ours):
            assert rel_l2(b, a) < 1e-6, step
    assert torch.equal(ours[4].detach(), ref[4].detach())


def test_trainer_labels_equal_onehots_fp32(cuda_dev):
    """GANTrainer.step on the uint8 wire format == on fp32 one-hots (parity mode, same eps stream): the H2D bytes drop 20x,
    the losses and the updated weights do not move."""
    res = []
    with dg.precision("fp32"):
        for labels in (False, True):
            torch.manual_seed(0)
            G = dg.Generator("relu", 9, 5, 13, 0.0, dim=128, depth=2, heads=8, mlp_ratio=3).to(cuda_dev)
            D = dg.Discriminator("relu", 9, 5, 13, 0.0, dim=128, depth=2, heads=8, mlp_ratio=3).to(cuda_dev)
            tr = gan.GANTrainer(G, D, lr_g=1e-3, lr_d=1e-3)
            mol = gan.synthetic_molecules(6, 9, 13, 5, seed=7, labels=labels, device=cuda_dev)
            drug = gan.synthetic_molecules(6, 9, 13, 5, seed=8, labels=labels, device=cuda_dev)
            torch.manual_seed(5)
            losses = [tr.step(drug[0], drug[1], mol[0], mol[1]) for _ in range(2)]
            res.append((losses, [p.detach().clone() for p in list(G.parameters()) + list(D.parameters())]))
    for l0, l1 in zip(res[0][0], res[1][0]):
        assert abs(l0[0] - l1[0]) < 1e-4 * max(1.0, abs(l0[0])) and abs(l0[1] - l1[1]) < 1e-4 * max(1.0, abs(l0[1]))
    for p0, p1 in zip(res[0][1], res[1][1]):
        assert rel_l2(p1, p0) < 1e-4


@pytest.mark.parametrize("B,N,D", [(1, 1, 128), (3, 9, 128), (2, 45, 128), (2, 7, 64)])
def test_symmetrize_bit_exact_and_self_adjoint(cuda_dev, B, N, D):
    """models.py:94 on a dense tensor: same two roundings as torch's add + div; the primitive's backward is the same launch."""
    e = torch.randn(B, N, N, D, generator=torch.Generator().manual_seed(N)).to(cuda_dev)
    want = (e + e.permute(0, 2, 1, 3)) / 2
    assert torch.equal(K.symmetrize(e), want)
    x = e.clone().requires_grad_(True)
    w = torch.randn_like(e)
    (ops.Symmetrize.apply(x) * w).sum().backward()
    assert torch.equal(x.grad, (w + w.permute(0, 2, 1, 3)) / 2)


def test_dense_prologue_matches_torch_modules(cuda_dev):
    """Discriminator on dense (generated-like) inputs: the fused prologue path (two Linear+ReLU primitives + one symmetrise pass)
    against the nn.Sequential + permute/add/div statement of models.py:92-94 in the fp32 parity mode, outputs and input gradients."""
    torch.manual_seed(0)
    D = dg.Discriminator("relu", 9, 5, 13, 0.0, dim=128, depth=1, heads=8, mlp_ratio=3).to(cuda_dev)
    z_e = torch.rand(4, 9, 9, 5, device=cuda_dev).requires_grad_(True)
    z_n = torch.rand(4, 9, 13, device=cuda_dev)
    with dg.precision("fp32"):
        out = D(z_e, z_n)
        (g1,) = torch.autograd.grad(out.sum(), z_e)
        edge = D.edge_layers(z_e)
        edge = (edge + edge.permute(0, 2, 1, 3)) / 2
        ref = D.node_mlp(D.TransformerEncoder(D.node_layers(z_n), edge)[0].reshape(4, -1))
        (g2,) = torch.autograd.grad(ref.sum(), z_e)
    assert rel_l2(out, ref) < 1e-5 and rel_l2(g1, g2) < 1e-4
