"""The fused tcgen05 edge-attention chain (dg_attn_edge_fwd), the bf16-score softmax-aggregate, bf16 `de` storage and
the L2-prefetch option, each against its plain-torch statement in fp64 on bf16-rounded operands (the tensor-core
kernels round contraction operands to bf16; everything else is fp32)."""
import os

import pytest
import torch

from druggen_b200 import _lib
from druggen_b200 import kernels as K
from druggen_b200.block import BLOCK_PARAM_NAMES, encoder_block
from emul_kernels import EmulBackend
from conftest import rel_l2

pytestmark = pytest.mark.gpu
EM = EmulBackend()


def rnd(dev, *shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed + len(shape) * 1000 + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(dev)


def bf(t):
    return t.to(torch.bfloat16).double()


def _ln(z, gamma, beta):
    mu, var = z.mean(-1, keepdim=True), z.var(-1, unbiased=False, keepdim=True)
    return (z - mu) / torch.sqrt(var + 1e-5) * gamma.double() + beta.double()


@pytest.mark.parametrize("B,N", [(1, 4), (3, 9), (2, 45), (40, 45), (300, 9), (5, 90), (148 * 3 + 1, 8)])
@pytest.mark.parametrize("side", ["none", "a16", "all"])
def test_attn_edge_fwd(cuda_dev, B, N, side):
    D, c = 128, 0.25
    R = B * N * N
    y = rnd(cuda_dev, R, D)
    q, k = rnd(cuda_dev, B, N, D, seed=1), rnd(cuda_dev, B, N, D, seed=2)
    we, be = rnd(cuda_dev, D, D, seed=3, scale=D ** -0.5), rnd(cuda_dev, D, seed=4, scale=0.1)
    woe, boe = rnd(cuda_dev, D, D, seed=5, scale=D ** -0.5), rnd(cuda_dev, D, seed=6, scale=0.1)
    gamma, beta = rnd(cuda_dev, D, seed=7, scale=0.1) + 1.0, rnd(cuda_dev, D, seed=8, scale=0.1)
    with K.precision("bf16"):
        out, a16, e, z = K.attn_edge_fwd(y, q, k, we, be, woe, boe, gamma, beta, c, want_a16=side != "none",
                                         want_e=side == "all", want_z=side == "all")
    e_ref = bf(y) @ bf(we).t() + be.double()
    qi = q.double().view(B, N, 1, D)
    kj = k.double().view(B, 1, N, D)
    e4 = e_ref.view(B, N, N, D)
    a_ref = (c * qi * kj * (e4 * e4 + e4)).reshape(R, D)
    if side == "none":
        assert a16 is None and e is None and z is None
        a_op = bf(a_ref.float())
    else:
        assert a16.dtype == torch.bfloat16
        assert rel_l2(a16.double(), bf(a_ref.float())) < 4e-3          # 1-ulp bf16 flips vs the fp64 emulation
        a_op = a16.double()                                             # the kernel's own operand isolates GEMM2 + LN
    z_ref = y.double() + a_op @ bf(woe).t() + boe.double()
    tol = 2e-4 if side != "none" else 3e-3
    assert rel_l2(out, _ln(z_ref, gamma, beta)) < tol, rel_l2(out, _ln(z_ref, gamma, beta))
    if side == "all":
        assert rel_l2(e, e_ref) < 2e-5
        assert rel_l2(z, z_ref) < 2e-5


@pytest.mark.parametrize("B,N", [(1, 4), (3, 9), (2, 45), (300, 9), (5, 90), (40, 45)])
def test_softmax_agg16(cuda_dev, B, N):
    D = 128
    a16 = rnd(cuda_dev, B, N, N, D, seed=3, scale=2.0).to(torch.bfloat16)
    v = rnd(cuda_dev, B, N, D, seed=2)
    a = a16.double()
    p = torch.softmax(a, dim=2)
    g_ref = (p * v.double().view(B, 1, N, D)).sum(2)
    g, stats = K.softmax_agg16_fwd(a16.view(-1, D), v, want_stats=True)
    assert rel_l2(g, g_ref) < 1e-5
    m = a.max(dim=2).values
    assert rel_l2(stats[0], m) < 1e-6
    assert rel_l2(stats[1], 1.0 / torch.exp(a - m[:, :, None, :]).sum(2)) < 1e-5
    assert rel_l2(K.softmax_agg16_fwd(a16.view(-1, D), v), g_ref) < 1e-5


@pytest.mark.parametrize("B,N", [(3, 9), (2, 45), (40, 45)])
def test_attn_scores_stats_only_and_bf16_de(cuda_dev, B, N):
    D, c = 128, 0.25
    q, k, v = rnd(cuda_dev, B, N, D), rnd(cuda_dev, B, N, D, seed=1), rnd(cuda_dev, B, N, D, seed=2)
    e, dg, da_in = rnd(cuda_dev, B, N, N, D, seed=3), rnd(cuda_dev, B, N, D, seed=4), rnd(cuda_dev, B, N, N, D, seed=5)
    a, g, stats = K.attn_scores_fwd(q, k, v, e, c, want_stats=True)
    a2, g2, stats2 = K.attn_scores_fwd(q, k, v, e, c, want_stats=True, store_a=False)
    # (the stats-only variant is the warp-per-query-atom kernel: same values up to the summation order)
    assert a2 is None and torch.equal(stats[0], stats2[0]) and rel_l2(g2, g) < 1e-6 and rel_l2(stats2[1], stats[1]) < 1e-6
    de, dq, dk, dv = K.attn_scores_bwd(dg, da_in, q, k, v, e, c, stats)
    de16, dq2, dk2, dv2 = K.attn_scores_bwd(dg, da_in, q, k, v, e, c, stats, de_bf16=True)
    assert de16.dtype == torch.bfloat16 and torch.equal(de16, de.to(torch.bfloat16))
    # (dq: four per-warp partial sums reduced with atomics when the statistics are given -- order-dependent in the last bit)
    assert rel_l2(dq2, dq) < 1e-6 and rel_l2(dk2, dk) < 1e-6 and rel_l2(dv2, dv) < 1e-6
    # the out_e path's gradient stored as bf16 (what rows_gemm(out_bf16=True) hands over in the tensor-core mode)
    da16 = da_in.to(torch.bfloat16)
    got16 = K.attn_scores_bwd(dg, da16, q, k, v, e, c, stats)
    want16 = K.attn_scores_bwd(dg, da16.float(), q, k, v, e, c, stats)
    for x_, w_ in zip(got16, want16):
        assert rel_l2(x_, w_) < 1e-6
    # statistics taken from bf16-stored scores (the chain path): the kernel rounds its recomputed scores alike
    a16 = a.to(torch.bfloat16)
    g16, st16 = K.softmax_agg16_fwd(a16.view(-1, D), v, want_stats=True)
    got = K.attn_scores_bwd(dg, da_in, q, k, v, e, c, st16, scores_bf16=True)
    d = lambda t: t.double()  # noqa: E731
    want = [torch.empty_like(d(e)), torch.empty_like(d(q)), torch.zeros_like(d(q)), torch.zeros_like(d(q))]
    EM.attn_scores_bwd(d(dg), d(da_in), d(q), d(k), d(v), d(e), c, want[0], want[1], want[2], want[3], None, True)
    for x_, w_ in zip(got, want):
        assert rel_l2(x_, w_) < 2e-3      # (1-ulp bf16 flips of individual scores between the fp32 kernel and the fp64 emulation)


def test_l2_prefetch_option_is_numerically_inert(cuda_dev):
    """DG_OPT_L2_PREFETCH only moves data into L2 early: results with and without it are identical."""
    R = 128 * 148 * 2 + 77
    x, dout = rnd(cuda_dev, R, 128), rnd(cuda_dev, R, 128, seed=7)
    w1, b1 = rnd(cuda_dev, 384, 128, seed=1, scale=128 ** -0.5), rnd(cuda_dev, 384, seed=2, scale=0.1)
    w2, b2 = rnd(cuda_dev, 128, 384, seed=3, scale=384 ** -0.5), rnd(cuda_dev, 128, seed=4, scale=0.1)
    gamma, beta = rnd(cuda_dev, 128, seed=5, scale=0.1) + 1.0, rnd(cuda_dev, 128, seed=6, scale=0.1)
    B, N = 37, 45
    q, k, v = rnd(cuda_dev, B, N, 128), rnd(cuda_dev, B, N, 128, seed=1), rnd(cuda_dev, B, N, 128, seed=2)
    e, dg, da_in = rnd(cuda_dev, B, N, N, 128, seed=3), rnd(cuda_dev, B, N, 128, seed=4), rnd(cuda_dev, B, N, N, 128, seed=5)

    def run():
        with K.precision("bf16"):
            o = [K.mlp_fwd(x, w1, b1, w2, b2, gamma, beta)]
            dz, h16, dgam, dbet = K.mlp_bwd_ln(x, dout, w1, b1, w2, b2, gamma)
            dx, dh16 = K.mlp_bwd_dgrad(dz, h16, w1, w2)
            o += [dz, h16.float(), dx, dh16.float()]
            o.append(K.rows_gemm(x, w1, True, b1, relu=True))
            o.append(K.rows_gemm(x, w2[:, :128].contiguous(), True, None, False, gate=dout, resid=dout))
            o.append(K.gemm_tn(x, dout))
            a, g, st = K.attn_scores_fwd(q, k, v, e, 0.25, want_stats=True)
            o += [a, g] + list(K.attn_scores_bwd(dg, da_in, q, k, v, e, 0.25, st))
            o += [t.float() for t in K.attn_edge_fwd(e.view(-1, 128), q, k, w1[:128].contiguous(), b1[:128].contiguous(),
                                                     w2[:, :128].contiguous(), b2, gamma, beta, 0.25, True, True, True)]
        return o
    try:
        K.set_option(_lib.OPT_L2_PREFETCH, 0)
        off = run()
        K.set_option(_lib.OPT_L2_PREFETCH, _lib.PF_ALL)
        on = run()
    finally:
        K.set_option(_lib.OPT_L2_PREFETCH, _lib.PF_DEFAULT)
    for i, (a, b) in enumerate(zip(off, on)):
        assert rel_l2(a, b) < 1e-6, i          # (atomically accumulated outputs differ in summation order only)


def _block_params(dev, d=128, r=3, seed=11):
    g = torch.Generator().manual_seed(seed)
    shapes = {"weight2": None}
    out = []
    for name in BLOCK_PARAM_NAMES:
        if name.startswith("ln"):
            t = torch.ones(d) + 0.1 * torch.randn(d, generator=g) if name.endswith("weight") else 0.1 * torch.randn(d, generator=g)
        elif "fc1" in name:
            t = torch.randn(r * d, d, generator=g) * d ** -0.5 if name.endswith("weight") else 0.1 * torch.randn(r * d, generator=g)
        elif "fc2" in name:
            t = torch.randn(d, r * d, generator=g) * (r * d) ** -0.5 if name.endswith("weight") else 0.1 * torch.randn(d, generator=g)
        else:
            t = torch.randn(d, d, generator=g) * d ** -0.5 if name.endswith("weight") else 0.1 * torch.randn(d, generator=g)
        out.append(t.to(dev))
    del shapes
    return out


@pytest.mark.parametrize("scores", ["bf16", "fp32"])
@pytest.mark.parametrize("B,N", [(3, 9), (4, 45)])
def test_block_chain_vs_unfused_attention(cuda_dev, B, N, scores):
    """The checkpointed block in the throughput mode with and without the fused edge-attention chain: same function,
    different rounding points (bf16 scores feed the forward's softmax) -> outputs and all gradients agree to bf16 level."""
    d, heads = 128, 8
    params = _block_params(cuda_dev)
    x0, y0 = rnd(cuda_dev, B, N, d, seed=1), rnd(cuda_dev, B, N, N, d, seed=2)
    wx, wy = rnd(cuda_dev, B, N, d, seed=3), rnd(cuda_dev, B, N, N, d, seed=4)

    def run(chain):
        os.environ["DRUGGEN_B200_ATTN_CHAIN"] = "1" if chain else "0"
        os.environ["DRUGGEN_B200_SOFTMAX_SCORES"] = scores
        try:
            with K.precision("bf16"):
                x, y = x0.clone().requires_grad_(True), y0.clone().requires_grad_(True)
                pp = [p.detach().clone().requires_grad_(True) for p in params]
                xo, yo = encoder_block(x, y, pp, heads, True)
                ((xo * wx).sum() + (yo * wy).sum()).backward()
                with torch.no_grad():
                    xn, yn = encoder_block(x0, y0, params, heads, True)
            return [xo.detach(), yo.detach(), xn, yn, x.grad, y.grad] + [p.grad for p in pp]
        finally:
            os.environ.pop("DRUGGEN_B200_ATTN_CHAIN", None)
            os.environ.pop("DRUGGEN_B200_SOFTMAX_SCORES", None)
    ref, got = run(False), run(True)
    # N(0,1) edge inputs make heavy-tailed scores (|a| up to ~20): with the softmax fed by bf16-stored scores each
    # probability carries |a| * 2^-9 relative error -- the bound is looser there than with fp32 scores
    tol = 4e-2 if scores == "bf16" else 1.5e-2
    for i, (a, b) in enumerate(zip(got, ref)):
        assert rel_l2(a, b) < tol, (i, scores, rel_l2(a, b))


@pytest.mark.gpu
@pytest.mark.parametrize("B,N", [(1, 4), (3, 9), (2, 45), (300, 45), (150, 17), (5, 48)])
def test_attn_scores_bwd_ring_equals_four_warp_kernel(cuda_dev, B, N):
    """dg_attn_scores_bwd with the forward's statistics runs the TMA-fed ring kernel (persistent CTAs, a warp owns <= 6 key
    atoms of every molecule, rows land in per-warp rings through cp.async.bulk); DG_OPT_ATTN_BWD = 1 forces the 4-warp
    register-staged kernel.  Same arithmetic per row; dq / dk / dv differ in the summation order only.  Every storage variant:
    da fp32 / bf16 / absent, de fp32 / bf16 / accumulated, scores rounded to bf16."""
    D, c = 128, 0.25
    q, k, v = rnd(cuda_dev, B, N, D), rnd(cuda_dev, B, N, D, seed=1), rnd(cuda_dev, B, N, D, seed=2)
    e, dg, da_in = rnd(cuda_dev, B, N, N, D, seed=3), rnd(cuda_dev, B, N, D, seed=4), rnd(cuda_dev, B, N, N, D, seed=5)
    _, _, stats = K.attn_scores_fwd(q, k, v, e, c, want_stats=True)
    acc0 = rnd(cuda_dev, B, N, N, D, seed=6)
    variants = [dict(da=da_in), dict(da=None), dict(da=da_in.to(torch.bfloat16), de_bf16=True, scores_bf16=True),
                dict(da=da_in, de_bf16=True), dict(da=da_in, accum=True)]

    def run(opt):
        K.set_option(_lib.OPT_ATTN_BWD, opt)
        out = []
        for vr in variants:
            acc = acc0.clone() if vr.get("accum") else None
            out.append(K.attn_scores_bwd(dg, vr["da"], q, k, v, e, c, stats, de_bf16=vr.get("de_bf16", False),
                                         scores_bf16=vr.get("scores_bf16", False), de_accum=acc))
        return out
    try:
        old = run(1)
        new = run(0)
        new16 = run(2)                                              # (the 16-warp form of the ring kernel)
    finally:
        K.set_option(_lib.OPT_ATTN_BWD, 0)
    for vi, (o, n_) in enumerate(list(zip(old, new)) + list(zip(old, new16))):
        # per-row arithmetic is the same expression (the compiler may contract its FMAs differently in the two kernels)
        assert rel_l2(n_[0].float(), o[0].float()) < (2e-4 if n_[0].dtype == torch.bfloat16 else 1e-6), ("de", vi)
        for nm, a_, b_ in zip(("dq", "dk", "dv"), n_[1:], o[1:]):
            assert rel_l2(a_, b_) < 2e-6, (nm, vi, rel_l2(a_, b_))



@pytest.mark.gpu
@pytest.mark.parametrize("B,N", [(3, 9), (2, 45)])
def test_kept_intermediates_on_the_gpu(cuda_dev, B, N):
    """block.keep_intermediates() in the throughput mode on the device: the block's backward takes what the forward kept (the
    fused edge chain is not launched a second time) and lands on the gradients of the recomputing path -- up to the order of the
    kernels' atomic reductions (dq, weight-gradient flushes), which also separates two runs of the same path."""
    from druggen_b200 import block as blk
    d, heads = 128, 8
    g = torch.Generator().manual_seed(21)
    mk = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(cuda_dev)  # noqa: E731
    params = []
    for nme in BLOCK_PARAM_NAMES:
        if nme.startswith("ln"):
            params.append(mk(d, sc=0.1) + (1.0 if nme.endswith("weight") else 0.0))
        elif "fc1.weight" in nme:
            params.append(mk(3 * d, d, sc=d ** -0.5))
        elif "fc1.bias" in nme:
            params.append(mk(3 * d, sc=0.1))
        elif "fc2.weight" in nme:
            params.append(mk(d, 3 * d, sc=(3 * d) ** -0.5))
        elif nme.endswith("weight"):
            params.append(mk(d, d, sc=d ** -0.5))
        else:
            params.append(mk(d, sc=0.1))
    x0, y0, wx, wy = mk(B, N, d), mk(B, N, N, d), mk(B, N, d), mk(B, N, N, d)
    be = K._be()

    def run(keep):
        l0 = be.launches
        x, y = x0.clone().requires_grad_(True), y0.clone().requires_grad_(True)
        pp = [p.clone().requires_grad_(True) for p in params]
        with K.precision("bf16"), blk.keep_intermediates(keep):
            xo, yo = encoder_block(x, y, pp, heads, True)
            ((xo * wx).sum() + (yo * wy).sum()).backward()
        torch.cuda.synchronize()
        return be.launches - l0, [xo.detach(), yo.detach(), x.grad, y.grad] + [p.grad for p in pp]
    n_re, ref = run(False)
    n_keep, got = run(True)
    # the recomputing backward re-launches LN1, q / k / v, the fused edge chain, out_n and LN3 (the softmax statistics are kept
    # either way): 7 launches that the keeping backward does not issue
    assert n_re - n_keep == 7, (n_re, n_keep)
    assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1])          # forward: the same launches
    for i, (a_, b_) in enumerate(zip(got[2:], ref[2:])):
        assert rel_l2(a_, b_) < 1e-4, (i, rel_l2(a_, b_))      # (the order of the fp32 atomic reductions moves x.grad by up to ~2e-5)
