"""Host logic (block composition, checkpointed block Functions, module API) on the torch
emulation of the kernel table, checked against the oracle and the reference golden vectors.
CPU only; the CUDA kernels themselves are covered by the -m gpu tests."""
import pytest
import torch

import druggen_b200 as dg
from druggen_b200 import kernels
from druggen_b200.block import BLOCK_PARAM_NAMES, block_forward, encoder_block
from conftest import load_golden, rel_l2, state_from
from emul_kernels import EmulBackend
from oracle import encoder_oracle as orc


@pytest.fixture(autouse=True)
def emul():
    kernels._install_backend_for_tests(EmulBackend())
    old = kernels.get_precision()
    kernels.set_precision("fp32")          # the emulation computes exactly; bf16 storage would only add rounding
    yield
    kernels.set_precision(old)
    kernels._install_backend_for_tests(None)


def _block_params(dtype=torch.float64, seed=0, d=128, r=3):
    g = torch.Generator().manual_seed(seed)
    p = {}
    for n in BLOCK_PARAM_NAMES:
        if n.startswith("ln"):
            shape, scale, shift = (d,), 0.1, (1.0 if n.endswith("weight") else 0.0)
        elif "fc1.weight" in n:
            shape, scale, shift = (r * d, d), d ** -0.5, 0.0
        elif "fc1.bias" in n:
            shape, scale, shift = (r * d,), 0.1, 0.0
        elif "fc2.weight" in n:
            shape, scale, shift = (d, r * d), (r * d) ** -0.5, 0.0
        elif n.endswith("weight"):
            shape, scale, shift = (d, d), d ** -0.5, 0.0
        else:
            shape, scale, shift = (d,), 0.1, 0.0
        p[n] = (torch.randn(shape, generator=g, dtype=dtype) * scale + shift).requires_grad_(True)
    return p


@pytest.mark.parametrize("edge_out", [True, False])
def test_block_first_and_second_order_vs_oracle(edge_out):
    """block_forward and the recompute-checkpointed EncoderBlockFn give the oracle's outputs,
    gradients and gradient-penalty-style second-order gradients (fp64)."""
    d, n, b, heads = 32, 4, 2, 4
    p = _block_params(d=d)
    g = torch.Generator().manual_seed(5)
    x0 = torch.randn(b, n, d, generator=g, dtype=torch.float64)
    y0 = torch.randn(b, n, n, d, generator=g, dtype=torch.float64)
    wx = torch.randn(b, n, d, generator=g, dtype=torch.float64)
    wy = torch.randn(b, n, n, d, generator=g, dtype=torch.float64)

    def run(kind):
        x, y = x0.clone().requires_grad_(True), y0.clone().requires_grad_(True)
        pp = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
        if kind == "oracle":
            xo, yo = orc.block_forward(x, y, {"blk." + k: v for k, v in pp.items()}, "blk.", heads)
        elif kind == "prim":
            xo, yo = block_forward(x, y, [pp[k] for k in BLOCK_PARAM_NAMES], heads, edge_out)
        else:
            xo, yo = encoder_block(x, y, [pp[k] for k in BLOCK_PARAM_NAMES], heads, edge_out)
        out = (xo * wx).sum() + ((yo * wy).sum() if edge_out else 0.0)
        gx, gy = torch.autograd.grad(out, [x, y], create_graph=True)
        pen = ((torch.cat([gx.reshape(b, -1), gy.reshape(b, -1)], 1).norm(2, dim=1) - 1) ** 2).mean()
        total = out + 3.0 * pen
        total.backward()
        return xo, yo, gx, gy, x.grad, y.grad, {k: v.grad for k, v in pp.items()}

    ref = run("oracle")
    for kind in ("prim", "ckpt"):
        got = run(kind)
        assert rel_l2(got[0], ref[0]) < 1e-10
        if edge_out:
            assert rel_l2(got[1], ref[1]) < 1e-10
        for i in (2, 3, 4, 5):
            assert rel_l2(got[i], ref[i]) < 1e-9, (kind, i)
        for k in BLOCK_PARAM_NAMES:
            if got[6][k] is None:      # parameter with no consumer (edge_out=False): reference grad is 0
                assert not edge_out and (ref[6][k] is None or float(ref[6][k].abs().max()) == 0.0), k
            else:
                assert rel_l2(got[6][k], ref[6][k]) < 1e-8, (kind, k)


def test_encoder_forward_golden_cfg1():
    g = load_golden("enc_fwd_cfg1.npz")
    enc = dg.TransformerEncoder(dim=128, depth=1, heads=8, act=None, mlp_ratio=3, drop_rate=0.0)
    enc.load_state_dict(state_from(g, "w::"))          # reference state-dict keys load unchanged
    with torch.no_grad():
        xo, yo = enc(torch.from_numpy(g["x"]), torch.from_numpy(g["y"]))
    assert rel_l2(xo, g["x_out"]) < 2e-5 and rel_l2(yo, g["y_out"]) < 2e-5
    assert yo.is_contiguous() and xo.is_contiguous()


def test_encoder_grads_golden():
    g = load_golden("enc_grad.npz")
    enc = dg.TransformerEncoder(dim=128, depth=2, heads=4, act=None, mlp_ratio=3, drop_rate=0.0)
    enc.load_state_dict(state_from(g, "w::"))
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    y = torch.from_numpy(g["y"]).requires_grad_(True)
    xo, yo = enc(x, y)
    ((xo * torch.from_numpy(g["wx"])).sum() + (yo * torch.from_numpy(g["wy"])).sum()).backward()
    assert rel_l2(x.grad, g["dx"]) < 1e-4 and rel_l2(y.grad, g["dy"]) < 1e-4
    for k, v in enc.named_parameters():
        assert rel_l2(v.grad, g["g::" + k]) < 1e-4, k


def _models(g):
    n, m_dim, b_dim = int(g["n"]), int(g["m_dim"]), int(g["b_dim"])
    G = dg.Generator("relu", n, b_dim, m_dim, 0.0, dim=128, depth=int(g["depth"]), heads=int(g["heads"]), mlp_ratio=3)
    D = dg.Discriminator("relu", n, b_dim, m_dim, 0.0, dim=128, depth=int(g["depth"]), heads=int(g["heads"]), mlp_ratio=3)
    G.load_state_dict(state_from(g, "wG::"))
    D.load_state_dict(state_from(g, "wD::"))
    return G, D


def test_gan_losses_golden_with_unchanged_loss_logic():
    """The GAN step of train.py:351-384 (losses restated in oracle, run ON OUR modules) reproduces
    the reference's losses and every gradient, including the gradient-penalty double backward and
    the None-gradient set of the Discriminator's dead last-block edge parameters."""
    g = load_golden("gan_step.npz")
    G, D = _models(g)
    t = {k: torch.from_numpy(g[k]) for k in ("drug_a", "drug_x", "mol_a", "mol_x", "eps_edge", "eps_node")}
    d_loss = orc.discriminator_loss(G, D, t["drug_a"], t["drug_x"], t["mol_a"], t["mol_x"],
                                    t["eps_edge"], t["eps_node"], float(g["lambda_gp"]))
    d_loss.backward()
    assert abs(d_loss.item() - float(g["d_loss"])) < 1e-4 * max(1.0, abs(float(g["d_loss"])))
    for k, v in D.named_parameters():
        gold = g["gD_D::" + k]
        if v.grad is None:
            assert float(abs(gold).max()) == 0.0, k
        else:
            assert rel_l2(v.grad, gold) < 2e-3, k
    assert all(p.grad is None for p in G.parameters())   # fake samples are detached (loss.py:62)
    D.zero_grad(set_to_none=True)
    g_loss = orc.generator_loss(G, D, t["mol_a"], t["mol_x"])
    g_loss.backward()
    assert abs(g_loss.item() - float(g["g_loss"])) < 1e-4 * max(1.0, abs(float(g["g_loss"])))
    for k, v in G.named_parameters():
        assert rel_l2(v.grad, g["gG_G::" + k]) < 1e-3, k
    dead = [k for k, v in D.named_parameters() if v.grad is None]
    assert sorted(dead) == sorted(k for k in dict(D.named_parameters())
                                  if float(abs(g["gG_D::" + k]).max()) == 0.0)


def test_decode_bit_exact_away_from_ties():
    g = load_golden("gan_step.npz")
    G, _ = _models(g)
    with torch.inference_mode():
        _, _, ns, es = G(torch.from_numpy(g["mol_a"]), torch.from_numpy(g["mol_x"]))
    safe_n, safe_e = torch.from_numpy(g["node_gap"]) > 1e-4, torch.from_numpy(g["edge_gap"]) > 1e-4
    assert torch.equal(ns.argmax(-1)[safe_n], torch.from_numpy(g["node_argmax"])[safe_n])
    assert torch.equal(es.argmax(-1)[safe_e], torch.from_numpy(g["edge_argmax"])[safe_e])


def test_product_refuses_cpu_without_test_backend():
    kernels._install_backend_for_tests(None)
    with pytest.raises(RuntimeError):
        dg.TransformerEncoder(128, 1, 8, None, 3, 0.0)(torch.zeros(1, 2, 128), torch.zeros(1, 2, 2, 128))


def test_training_dropout_runs_and_matches_reference_semantics():
    """--dropout / --ddropout > 0 in training mode (reference default 0): the block applies torch dropout to the two MLP outputs
    (layers.py:54), nothing else.  p > 0 in train() differs from eval(), is differentiable, and with p -> 0 equals eval()."""
    torch.manual_seed(0)
    enc = dg.TransformerEncoder(32, 2, 4, None, 3, 0.5).double()
    x = torch.randn(2, 4, 32, dtype=torch.float64, requires_grad=True)
    y = torch.randn(2, 4, 4, 32, dtype=torch.float64, requires_grad=True)
    xo_t, yo_t = enc(x, y)                                  # train(): dropout active
    (xo_t.sum() + yo_t.sum()).backward()
    assert x.grad is not None and all(p.grad is not None for p in enc.parameters())
    enc.eval()
    xo_e, yo_e = enc(x, y)
    assert rel_l2(xo_t, xo_e) > 1e-3                        # the masks did something
    enc.train()
    for blk in enc.Encoder_Blocks:
        blk._drop = 1e-12                                   # keeps every element with probability ~1: the dropout path, no masking
    xo_p, yo_p = enc(x, y)
    assert rel_l2(xo_p, xo_e) < 1e-9 and rel_l2(yo_p, yo_e) < 1e-9


def test_fused_branch_wiring_in_throughput_mode():
    """With precision 'bf16' the block routes through the fused-kernel entry points (mlp_fwd, mlp_bwd_ln,
    mlp_bwd_dgrad, attn_scores_*, bf16-stored hidden).  On the emulation the only difference from the oracle is the
    bf16 rounding of the stored hidden activation, so gradients agree to ~1e-2."""
    kernels.set_precision("bf16")
    d, n, b, heads = 128, 5, 2, 8
    p = _block_params(dtype=torch.float32, d=d)
    g = torch.Generator().manual_seed(6)
    x0, y0 = torch.randn(b, n, d, generator=g), torch.randn(b, n, n, d, generator=g)
    wx, wy = torch.randn(b, n, d, generator=g), torch.randn(b, n, n, d, generator=g)

    def run(kind):
        x, y = x0.clone().requires_grad_(True), y0.clone().requires_grad_(True)
        pp = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
        if kind == "oracle":
            xo, yo = orc.block_forward(x, y, {"blk." + k: v for k, v in pp.items()}, "blk.", heads)
        else:
            xo, yo = encoder_block(x, y, [pp[k] for k in BLOCK_PARAM_NAMES], heads, True)
        ((xo * wx).sum() + (yo * wy).sum()).backward()
        return xo, yo, x.grad, y.grad, {k: v.grad for k, v in pp.items()}

    ref, got = run("oracle"), run("ckpt")
    for i in range(4):
        assert rel_l2(got[i], ref[i]) < 2e-2, i
    for k in BLOCK_PARAM_NAMES:
        assert rel_l2(got[4][k], ref[4][k]) < 3e-2, k


def test_mlp_primitive_chain_wiring():
    """In the throughput mode the residual MLP primitive runs its first- and second-order backward on the fused dgrad chain
    (second order with the two weights transposed into each other's role): same gradients as the unfused fp32 primitive up to
    the bf16 storage of h / dh."""
    from druggen_b200 import ops
    g = torch.Generator().manual_seed(3)
    x0 = torch.randn(40, 128, generator=g)
    w1, b1 = torch.randn(384, 128, generator=g) * 128 ** -0.5, torch.randn(384, generator=g) * 0.1
    w2, b2 = torch.randn(128, 384, generator=g) * 384 ** -0.5, torch.randn(128, generator=g) * 0.1
    wgt, u = torch.randn(40, 128, generator=g), torch.randn(40, 128, generator=g)

    def run(narrow):
        kernels.set_precision("bf16" if narrow else "fp32")
        x = x0.clone().requires_grad_(True)
        pp = [t.clone().requires_grad_(True) for t in (w1, b1, w2, b2)]
        m = ops.MLP.apply(x, pp[0], pp[1], pp[2], pp[3], narrow, True)
        (dx,) = torch.autograd.grad(m, x, wgt, create_graph=True)            # first order, with a graph (gradient penalty)
        second = torch.autograd.grad(dx, [x] + pp, u, allow_unused=True)      # second order w.r.t. everything
        return [m.detach(), dx.detach()] + [t for t in second]
    ref, got = run(False), run(True)
    for i, (a, b) in enumerate(zip(got, ref)):
        if b is None:
            assert a is None or float(a.abs().max()) == 0.0, i
        else:
            assert rel_l2(a, b) < 2e-2, (i, rel_l2(a, b))


def test_frozen_discriminator_in_g_step_changes_nothing_that_is_read():
    """GANTrainer(skip_dead_d_grads=True) freezes D while the G-step graph is built (train.py:371-377 computes D weight
    gradients that reset_grad, train.py:352, discards).  One full step with and without the freeze: identical losses,
    identical G and D weights afterwards; with the freeze D.grad is simply never populated by the G-step."""
    from druggen_b200 import gan

    def run(skip):
        torch.manual_seed(0)
        G = dg.Generator("relu", 5, 5, 13, 0.0, dim=32, depth=2, heads=4, mlp_ratio=3)
        D = dg.Discriminator("relu", 5, 5, 13, 0.0, dim=32, depth=2, heads=4, mlp_ratio=3)
        tr = gan.GANTrainer(G, D, lr_g=1e-3, lr_d=1e-3, skip_dead_d_grads=skip)
        a, x = gan.synthetic_molecules(4, 5, 13, 5, seed=7)
        da, dx = gan.synthetic_molecules(4, 5, 13, 5, seed=8)
        torch.manual_seed(5)                      # the gradient penalty's eps draws
        losses = tr.step(da, dx, a, x)
        d_grads = [p.grad is not None for p in D.parameters()]
        assert all(p.requires_grad for p in D.parameters())          # the freeze is undone when the G-step graph is built
        return losses, [p.detach().clone() for p in G.parameters()], [p.detach().clone() for p in D.parameters()], d_grads

    l0, g0, d0, dg0 = run(False)
    l1, g1, d1, dg1 = run(True)
    assert l0 == l1
    for a_, b_ in zip(g0 + d0, g1 + d1):
        assert torch.equal(a_, b_)
    assert any(dg0) and not any(dg1)              # reference behaviour leaves dead D grads behind; the freeze leaves none


def test_io_wrappers_on_the_emulation():
    """label2onehot / argmax_last host wrappers (shapes, dtypes, reference signature) on the CPU emulation"""
    labels = torch.randint(0, 5, (2, 4, 4))
    want = torch.zeros(2, 4, 4, 5).scatter_(3, labels.unsqueeze(-1), 1.0)
    assert torch.equal(dg.label2onehot(labels, 5), want)
    assert torch.equal(dg.label2onehot(labels.to(torch.uint8), 5), want)
    with pytest.raises(RuntimeError):
        dg.label2onehot(labels.float(), 5)
    t = torch.randn(3, 7, 5)
    assert torch.equal(dg.argmax_last(t), torch.max(t, -1)[1])


@pytest.mark.parametrize("mode,d,tol", [("fp32", 32, 1e-9), ("bf16", 128, 1e-6)])
@pytest.mark.parametrize("edge_out", [True, False])
def test_hand_sequenced_second_order_equals_autograd_path(monkeypatch, mode, d, tol, edge_out):
    """block_backward_backward (the hand-sequenced second-order pass the gradient penalty takes) against reverse-over-reverse
    on the differentiable primitives: same cotangents for x, y and every parameter.  The bf16 case runs the narrow code paths
    (bf16-stored h / dh, the dgrad chain in both weight roles, the fused edge-attention recompute) on the emulation."""
    from druggen_b200 import block as blk
    n, b, heads = 4, 2, 8
    p = _block_params(d=d)
    g = torch.Generator().manual_seed(7)
    x0 = torch.randn(b, n, d, generator=g, dtype=torch.float64)
    y0 = torch.randn(b, n, n, d, generator=g, dtype=torch.float64)
    wx = torch.randn(b, n, d, generator=g, dtype=torch.float64)
    wy = torch.randn(b, n, n, d, generator=g, dtype=torch.float64)
    calls = []
    real = blk.block_backward_backward
    monkeypatch.setattr(blk, "block_backward_backward", lambda *a, **k: (calls.append(1), real(*a, **k))[1])

    def run(hand):
        monkeypatch.setattr(blk, "_HAND_SECOND_ORDER", hand)
        x, y = x0.clone().requires_grad_(True), y0.clone().requires_grad_(True)
        pp = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
        xo, yo = encoder_block(x, y, [pp[k] for k in BLOCK_PARAM_NAMES], heads, edge_out)
        out = (xo * wx).sum() + ((yo * wy).sum() if edge_out else 0.0)
        gx, gy = torch.autograd.grad(out, [x, y], create_graph=True)
        pen = ((torch.cat([gx.reshape(b, -1), gy.reshape(b, -1)], 1).norm(2, dim=1) - 1) ** 2).mean()
        pen.backward()                     # the second-order terms alone
        return x.grad, y.grad, {k: v.grad for k, v in pp.items()}

    with kernels.precision(mode):
        ref = run(False)
        assert not calls
        got = run(True)
        assert calls
    assert rel_l2(got[0], ref[0]) < tol and rel_l2(got[1], ref[1]) < tol
    for k in BLOCK_PARAM_NAMES:
        assert (got[2][k] is None) == (ref[2][k] is None), k
        if ref[2][k] is not None and float(ref[2][k].abs().max()) > 0:
            assert rel_l2(got[2][k], ref[2][k]) < tol, k


def test_kept_intermediates_equal_recomputation():
    """block.keep_intermediates(): the checkpointed block keeps what its backward would recompute (the fused chain's outputs).
    Same launches, same inputs -> outputs and every gradient bit-identical to the recomputing path; and the kept path really
    skips the chain's second run."""
    from druggen_b200 import block as blk
    kernels.set_precision("bf16")
    d, n, b, heads = 128, 5, 2, 8
    p = _block_params(dtype=torch.float32, d=d)
    g = torch.Generator().manual_seed(16)
    x0, y0 = torch.randn(b, n, d, generator=g), torch.randn(b, n, n, d, generator=g)
    wx, wy = torch.randn(b, n, d, generator=g), torch.randn(b, n, n, d, generator=g)
    calls = []
    be = kernels._test_backend
    orig = be.attn_edge_fwd
    be.attn_edge_fwd = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]

    def run(keep):
        calls.clear()
        x, y = x0.clone().requires_grad_(True), y0.clone().requires_grad_(True)
        pp = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
        with blk.keep_intermediates(keep):
            xo, yo = encoder_block(x, y, [pp[k] for k in BLOCK_PARAM_NAMES], heads, True)
        ((xo * wx).sum() + (yo * wy).sum()).backward()
        return len(calls), [xo.detach(), yo.detach(), x.grad, y.grad] + [pp[k].grad for k in BLOCK_PARAM_NAMES]

    try:
        n_re, ref = run(False)
        n_keep, got = run(True)
    finally:
        be.attn_edge_fwd = orig
    assert (n_re, n_keep) == (2, 1)
    for i, (a_, b_) in enumerate(zip(got, ref)):
        assert torch.equal(a_, b_), i


@pytest.mark.parametrize("mode,dim", [("fp32", 32), ("bf16", 128)])
def test_sequenced_d_step_equals_single_backward(mode, dim):
    """GANTrainer(sequenced=True) backpropagates real, fake and the gradient penalty one after the other (and lets the plain
    passes keep their intermediates): same losses and the same Discriminator gradients at the optimizer step as train.py's
    single d_loss.backward(), up to the fp32 accumulation order of the three terms.  (The Generator step then sees D weights
    that differ in the last bits -- AdamW's first step divides by |g| -- so its gradients are compared loosely.)"""
    from druggen_b200 import gan
    kernels.set_precision(mode)

    def run(seq):
        torch.manual_seed(0)
        G = dg.Generator("relu", 5, 5, 13, 0.0, dim=dim, depth=2, heads=4, mlp_ratio=3)
        D = dg.Discriminator("relu", 5, 5, 13, 0.0, dim=dim, depth=2, heads=4, mlp_ratio=3)
        tr = gan.GANTrainer(G, D, lr_g=1e-3, lr_d=1e-3, sequenced=seq)
        rec = {}
        for nm, opt, net in (("d", tr.d_optimizer, D), ("g", tr.g_optimizer, G)):
            def step(orig=opt.step, nm=nm, net=net):
                rec[nm] = [None if p.grad is None else p.grad.detach().clone() for p in net.parameters()]
                return orig()
            opt.step = step
        a, x = gan.synthetic_molecules(4, 5, 13, 5, seed=7)
        da, dx = gan.synthetic_molecules(4, 5, 13, 5, seed=8)
        torch.manual_seed(5)                      # the gradient penalty's eps draws
        return tr.step(da, dx, a, x), rec

    (d0, g0), r0 = run(False)
    (d1, g1), r1 = run(True)
    assert abs(d0 - d1) <= 1e-5 * max(1.0, abs(d0)) and abs(g0 - g1) <= 1e-4 * max(1.0, abs(g0))
    for nm, tol in (("d", 1e-6), ("g", 2e-3)):
        for a_, b_ in zip(r0[nm], r1[nm]):
            assert (a_ is None) == (b_ is None)
            if a_ is not None:
                assert rel_l2(b_, a_) < tol, (nm, rel_l2(b_, a_))
