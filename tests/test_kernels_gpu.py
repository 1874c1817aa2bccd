"""Each CUDA kernel (through the C-ABI) against its plain-torch statement (tests/emul_kernels.py,
run on the same device in fp64 where it matters).  Covers ragged row counts, N in {1, 9, 45},
optional operands, and the atomically-accumulated outputs."""
import pytest
import torch

from druggen_b200 import kernels as K
from emul_kernels import EmulBackend
from conftest import rel_l2

pytestmark = pytest.mark.gpu
EM = EmulBackend()


def rnd(dev, *shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed + len(shape) * 1000 + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(dev)


def _opr(t, tc):
    """operand as the kernel sees it: bf16-rounded when the tcgen05 kernel takes the shape (other
    shapes -- channel counts that are not multiples of 64/128 -- run the fp32 CUDA-core kernel)"""
    return (t.to(torch.bfloat16) if tc else t).double()


@pytest.mark.parametrize("prec", ["fp32", "bf16", "bf16x3"])
@pytest.mark.parametrize("R,Kd,Nd", [(1, 128, 128), (257, 128, 384), (1000, 384, 128), (2025 * 3, 128, 128), (77, 32, 32),
                                     (128 * 148 * 3 + 5, 128, 128), (40000, 128, 384), (40000, 384, 128),
                                     (777, 128, 512), (777, 512, 128)])     # mlp_ratio = 4: 128-wide column slices
def test_rows_gemm(cuda_dev, prec, R, Kd, Nd):
    a, bias = rnd(cuda_dev, R, Kd), rnd(cuda_dev, Nd, seed=3)
    gate = rnd(cuda_dev, R, Nd, seed=4)
    with K.precision(prec):
        for w_is_nk in (True, False):
            w = rnd(cuda_dev, Nd, Kd, seed=1) if w_is_nk else rnd(cuda_dev, Kd, Nd, seed=1)
            tc = prec == "bf16" and Kd % 64 == 0 and Nd % 128 == 0
            ref = _opr(a, tc) @ (_opr(w, tc).t() if w_is_nk else _opr(w, tc))
            # bf16x3: operands carry 16 mantissa bits (hi + lo), the a_lo w_lo term is dropped: ~1e-5 on a K = 128..384 dot product
            tol = 3e-5 if prec == "bf16x3" else 2e-6
            assert rel_l2(K.rows_gemm(a, w, w_is_nk), ref) < tol
            got = K.rows_gemm(a, w, w_is_nk, bias, True, gate)
            want = torch.relu(ref + bias.double()) * (gate > 0)
            assert rel_l2(got, want) < tol
            if (prec == "bf16x3" and Kd > 128) or (prec == "bf16" and Kd > 384):    # (K-sliced launches: resid only with a linear epilogue)
                got = K.rows_gemm(a, w, w_is_nk, bias, resid=gate)
                assert rel_l2(got, ref + bias.double() + gate.double()) < tol
                continue
            got = K.rows_gemm(a, w, w_is_nk, None, False, gate, resid=gate)      # fused gradient accumulation
            assert rel_l2(got, ref * (gate > 0) + gate.double()) < tol


@pytest.mark.parametrize("prec", ["fp32", "bf16", "bf16x3"])
@pytest.mark.parametrize("R,M,N", [(1, 128, 128), (333, 128, 384), (2025 * 4 + 5, 384, 128), (50, 32, 64),
                                   (64 * 148 * 5 + 3, 128, 128), (100000, 128, 384), (100000, 384, 128),
                                   (999, 128, 512), (999, 512, 128)])       # mlp_ratio = 4: 128 x 128 output blocks
def test_gemm_tn(cuda_dev, prec, R, M, N):
    a, b = rnd(cuda_dev, R, M), rnd(cuda_dev, R, N, seed=1)
    with K.precision(prec):
        got = K.gemm_tn(a, b)
        tc = prec == "bf16" and M % 128 == 0 and N % 128 == 0
        ref = _opr(a, tc).t() @ _opr(b, tc)
        tol = 3e-5 if prec == "bf16x3" else 5e-6
        assert rel_l2(got, ref) < tol
        cs = torch.zeros(M, device=cuda_dev)
        acc = K.gemm_tn(a, b, out=got.clone(), colsum_a=cs)
        assert rel_l2(acc, 2 * ref) < tol
        assert rel_l2(cs, a.double().sum(0)) < 5e-6          # bias gradient from the same pass, exact fp32 inputs


def test_colsum_gate(cuda_dev):
    for R, N in [(1, 128), (4099, 384), (100, 64)]:
        a = rnd(cuda_dev, R, N)
        assert rel_l2(K.colsum(a), a.double().sum(0)) < 2e-6
    for n in [5, 128, 1000003]:
        x, r = rnd(cuda_dev, n), rnd(cuda_dev, n, seed=1)
        assert torch.equal(K.gate_mul(x, r), x * (r > 0))


@pytest.mark.parametrize("R,D", [(1, 128), (7, 128), (1031, 128), (65, 384), (33, 32)])
@pytest.mark.parametrize("has_b", [True, False])
def test_add_ln(cuda_dev, R, D, has_b):
    a, b = rnd(cuda_dev, R, D), (rnd(cuda_dev, R, D, seed=1) if has_b else None)
    gamma, beta = rnd(cuda_dev, D, seed=2) + 1.0, rnd(cuda_dev, D, seed=3)
    dy, u = rnd(cuda_dev, R, D, seed=4), rnd(cuda_dev, R, D, seed=5)
    vg, vb = rnd(cuda_dev, D, seed=6), rnd(cuda_dev, D, seed=7)
    d = lambda t: None if t is None else t.double()  # noqa: E731
    out = torch.empty_like(a, dtype=torch.float64)
    EM.add_ln_fwd(d(a), d(b), d(gamma), d(beta), out, 1e-5)
    assert rel_l2(K.add_ln_fwd(a, b, gamma, beta), out) < 2e-6
    dz, dgm, dbt = (torch.empty(R, D, dtype=torch.float64, device=cuda_dev), torch.empty(D, dtype=torch.float64, device=cuda_dev),
                    torch.empty(D, dtype=torch.float64, device=cuda_dev))
    EM.add_ln_bwd(d(dy), d(a), d(b), d(gamma), dz, dgm, dbt, 1e-5)
    got = K.add_ln_bwd(dy, a, b, gamma)
    for g_, w_ in zip(got, (dz, dgm, dbt)):
        assert rel_l2(g_, w_) < 5e-6
    for vgi, vbi in ((vg, vb), (None, None)):
        gdy, gz, gg = torch.empty_like(dz), torch.empty_like(dz), torch.empty_like(dgm)
        EM.add_ln_bwd_bwd(d(u), d(vgi), d(vbi), d(dy), d(a), d(b), d(gamma), gdy, gz, gg, 1e-5)
        got = K.add_ln_bwd_bwd(u, vgi, vbi, dy, a, b, gamma)
        for g_, w_ in zip(got, (gdy, gz, gg)):
            assert rel_l2(g_, w_) < 2e-5


@pytest.mark.parametrize("B,N,D", [(1, 1, 128), (3, 9, 128), (2, 45, 128), (300, 9, 128), (2, 5, 64)])
def test_modulate(cuda_dev, B, N, D):
    q, k, e = rnd(cuda_dev, B, N, D), rnd(cuda_dev, B, N, D, seed=1), rnd(cuda_dev, B, N, N, D, seed=2)
    da = rnd(cuda_dev, B, N, N, D, seed=3)
    uq, uk, ue = rnd(cuda_dev, B, N, D, seed=4), rnd(cuda_dev, B, N, D, seed=5), rnd(cuda_dev, B, N, N, D, seed=6)
    c = 0.25
    d = lambda t: t.double()  # noqa: E731
    out = torch.empty_like(e, dtype=torch.float64)
    EM.modulate_fwd(d(q), d(k), d(e), c, out)
    assert rel_l2(K.modulate_fwd(q, k, e, c), out) < 2e-6
    dq, dk, de = torch.empty_like(d(q)), torch.empty_like(d(k)), torch.empty_like(out)
    EM.modulate_bwd(d(da), d(q), d(k), d(e), c, dq, dk, de)
    for g_, w_ in zip(K.modulate_bwd(da, q, k, e, c), (dq, dk, de)):
        assert rel_l2(g_, w_) < 5e-6
    gda, gq, gk, ge = torch.empty_like(out), torch.empty_like(dq), torch.empty_like(dq), torch.empty_like(out)
    EM.modulate_bwd_bwd(d(uq), d(uk), d(ue), d(da), d(q), d(k), d(e), c, gda, gq, gk, ge)
    for g_, w_ in zip(K.modulate_bwd_bwd(uq, uk, ue, da, q, k, e, c), (gda, gq, gk, ge)):
        assert rel_l2(g_, w_) < 5e-6


@pytest.mark.parametrize("B,N,D", [(1, 1, 128), (3, 9, 128), (2, 45, 128), (300, 9, 128), (2, 5, 64)])
def test_softmax_agg(cuda_dev, B, N, D):
    a, v = rnd(cuda_dev, B, N, N, D, scale=3.0), rnd(cuda_dev, B, N, D, seed=1)
    dg = rnd(cuda_dev, B, N, D, seed=2)
    ua, uv = rnd(cuda_dev, B, N, N, D, seed=3), rnd(cuda_dev, B, N, D, seed=4)
    d = lambda t: t.double()  # noqa: E731
    g = torch.empty_like(d(v))
    EM.softmax_agg_fwd(d(a), d(v), g)
    assert rel_l2(K.softmax_agg_fwd(a, v), g) < 5e-6
    da, dv = torch.empty_like(d(a)), torch.empty_like(g)
    EM.softmax_agg_bwd(d(dg), d(a), d(v), da, dv)
    for g_, w_ in zip(K.softmax_agg_bwd(dg, a, v), (da, dv)):
        assert rel_l2(g_, w_) < 1e-5
    got_da, _ = K.softmax_agg_bwd(dg, a, v, da_accum=ua.clone())
    assert rel_l2(got_da, da + ua.double()) < 1e-5
    gdg, ga, gv = torch.empty_like(g), torch.empty_like(da), torch.empty_like(g)
    EM.softmax_agg_bwd_bwd(d(ua), d(uv), d(dg), d(a), d(v), gdg, ga, gv)
    for g_, w_ in zip(K.softmax_agg_bwd_bwd(ua, uv, dg, a, v), (gdg, ga, gv)):
        assert rel_l2(g_, w_) < 2e-5


def test_rejects_cpu_and_bad_shapes(cuda_dev):
    with pytest.raises(RuntimeError):
        K.add_ln_fwd(torch.zeros(2, 128), None, torch.ones(128), torch.zeros(128))
    with pytest.raises(RuntimeError):   # D not a multiple of 4
        K.add_ln_fwd(torch.zeros(2, 6, device=cuda_dev), None, torch.ones(6, device=cuda_dev), torch.zeros(6, device=cuda_dev))
    z = torch.zeros(0, 128, device=cuda_dev)   # empty inputs are legal no-ops
    assert K.add_ln_fwd(z, None, torch.ones(128, device=cuda_dev), torch.zeros(128, device=cuda_dev)).shape == (0, 128)


@pytest.mark.parametrize("R", [1, 127, 128, 129, 2025, 128 * 148 + 77, 128 * 148 * 3 + 5])
@pytest.mark.parametrize("H", [384, 128])
def test_fused_mlp_fwd(cuda_dev, R, H):
    """fused tcgen05 residual-MLP vs the same arithmetic in fp64 on bf16-rounded operands (x, W, hidden)."""
    x = rnd(cuda_dev, R, 128)
    w1, b1 = rnd(cuda_dev, H, 128, seed=1, scale=128 ** -0.5), rnd(cuda_dev, H, seed=2, scale=0.1)
    w2, b2 = rnd(cuda_dev, 128, H, seed=3, scale=H ** -0.5), rnd(cuda_dev, 128, seed=4, scale=0.1)
    gamma, beta = rnd(cuda_dev, 128, seed=5, scale=0.1) + 1.0, rnd(cuda_dev, 128, seed=6, scale=0.1)
    bf = lambda t: t.to(torch.bfloat16).double()  # noqa: E731
    h = torch.relu(bf(x) @ bf(w1).t() + b1.double())
    z = x.double() + bf(h.float()) @ bf(w2).t() + b2.double()
    mu, var = z.mean(-1, keepdim=True), z.var(-1, unbiased=False, keepdim=True)
    want = (z - mu) / torch.sqrt(var + 1e-5) * gamma.double() + beta.double()
    with K.precision("bf16"):
        got = K.mlp_fwd(x, w1, b1, w2, b2, gamma, beta)
    assert rel_l2(got, want) < 2e-4, rel_l2(got, want)   # bf16 rounding of the hidden can flip by 1 ulp vs fp64 emulation
    assert float((got.double() - want).abs().max()) < 5e-2


@pytest.mark.parametrize("B,N", [(1, 4), (3, 9), (2, 45), (300, 9), (5, 90), (40, 45)])
def test_fused_attn_scores(cuda_dev, B, N):
    D, c = 128, 0.25
    q, k, v = rnd(cuda_dev, B, N, D), rnd(cuda_dev, B, N, D, seed=1), rnd(cuda_dev, B, N, D, seed=2)
    e, dg, da_in = rnd(cuda_dev, B, N, N, D, seed=3), rnd(cuda_dev, B, N, D, seed=4), rnd(cuda_dev, B, N, N, D, seed=5)
    d = lambda t: t.double()  # noqa: E731
    a64, g64 = torch.empty_like(d(e)), torch.empty_like(d(q))
    EM.attn_scores_fwd(d(q), d(k), d(v), d(e), c, a64, g64)
    a, g, stats = K.attn_scores_fwd(q, k, v, e, c, want_stats=True)
    assert rel_l2(a, a64) < 2e-6 and rel_l2(g, g64) < 1e-5
    for din in (da_in, None):
        de64, dq64, dk64, dv64 = torch.empty_like(a64), torch.empty_like(g64), torch.empty_like(g64), torch.empty_like(g64)
        EM.attn_scores_bwd(d(dg), None if din is None else d(din), d(q), d(k), d(v), d(e), c, de64, dq64, dk64, dv64)
        for st in (None, stats):           # statistics recomputed in-kernel / taken from the forward
            for got, want in zip(K.attn_scores_bwd(dg, din, q, k, v, e, c, st), (de64, dq64, dk64, dv64)):
                assert rel_l2(got, want) < 2e-5


@pytest.mark.parametrize("R", [1, 300, 128 * 148 + 77])
def test_bf16_storage_gemms(cuda_dev, R):
    """bf16-stored hidden activation / gradient: relu->bf16 out, bf16 A operand, bf16 gate, bf16 wgrad inputs."""
    x, dz = rnd(cuda_dev, R, 128), rnd(cuda_dev, R, 128, seed=9)
    w1, b1, w2 = rnd(cuda_dev, 384, 128, seed=1, scale=0.1), rnd(cuda_dev, 384, seed=2, scale=0.1), rnd(cuda_dev, 128, 384, seed=3, scale=0.1)
    bf = lambda t: t.to(torch.bfloat16).double()  # noqa: E731
    with K.precision("bf16"):
        h16 = K.rows_gemm(x, w1, True, b1, relu=True, out_bf16=True)
        assert h16.dtype == torch.bfloat16
        h_ref = torch.relu(bf(x) @ bf(w1).t() + b1.double())
        assert rel_l2(h16.double(), bf(h_ref.float())) < 3e-3          # 1-ulp bf16 flips vs the fp64 emulation
        m = K.rows_gemm(h16, w2, True)                                  # bf16 A operand
        assert rel_l2(m, h16.double() @ bf(w2).t()) < 2e-5
        dh16 = K.rows_gemm(dz, w2, False, gate=h16, out_bf16=True)      # bf16 sign gate + bf16 out
        dh_ref = (bf(dz) @ bf(w2)) * (h16.double() > 0)
        assert rel_l2(dh16.double(), bf(dh_ref.float())) < 3e-3
        assert torch.equal(dh16 == 0, (h16 <= 0) | (dh16 == 0))
        gb = torch.zeros(384, device=cuda_dev)
        gw1 = K.gemm_tn(dh16, x, colsum_a=gb)                           # bf16 a (with fused bias grad), fp32 b
        assert rel_l2(gw1, dh16.double().t() @ bf(x)) < 2e-5
        assert rel_l2(gb, dh16.double().sum(0)) < 1e-5
        gw2 = K.gemm_tn(dz, h16)                                        # fp32 a, bf16 b
        assert rel_l2(gw2, bf(dz).t() @ h16.double()) < 2e-5


@pytest.mark.parametrize("R", [1, 129, 2025, 128 * 148 + 77, 128 * 148 * 2 + 5])
@pytest.mark.parametrize("H", [384, 128])
def test_fused_mlp_bwd(cuda_dev, R, H):
    """the two fused backward chains vs fp64 on bf16-rounded operands: LN backward + h spill, gated dgrad + dh spill"""
    x, dout = rnd(cuda_dev, R, 128), rnd(cuda_dev, R, 128, seed=7)
    w1, b1 = rnd(cuda_dev, H, 128, seed=1, scale=128 ** -0.5), rnd(cuda_dev, H, seed=2, scale=0.1)
    w2, b2 = rnd(cuda_dev, 128, H, seed=3, scale=H ** -0.5), rnd(cuda_dev, 128, seed=4, scale=0.1)
    gamma = rnd(cuda_dev, 128, seed=5, scale=0.1) + 1.0
    bf = lambda t: t.to(torch.bfloat16).double()  # noqa: E731
    with K.precision("bf16"):
        dz, h16, dgam, dbet = K.mlp_bwd_ln(x, dout, w1, b1, w2, b2, gamma)
        h_ref = torch.relu(bf(x) @ bf(w1).t() + b1.double())
        assert rel_l2(h16.double(), bf(h_ref.float())) < 3e-3
        z = x.double() + h16.double() @ bf(w2).t() + b2.double()          # from the kernel's own h: isolates the LN math
        mu, var = z.mean(-1, keepdim=True), z.var(-1, unbiased=False, keepdim=True)
        r = torch.rsqrt(var + 1e-5)
        xh = (z - mu) * r
        gh = dout.double() * gamma.double()
        dz_ref = r * (gh - gh.mean(-1, keepdim=True) - xh * (gh * xh).mean(-1, keepdim=True))
        assert rel_l2(dz, dz_ref) < 2e-4
        assert rel_l2(dgam, (dout.double() * xh).sum(0)) < 2e-4 and rel_l2(dbet, dout.double().sum(0)) < 1e-5
        dx, dh16 = K.mlp_bwd_dgrad(dz, h16, w1, w2)
        dh_ref = (bf(dz) @ bf(w2)) * (h16.double() > 0)
        assert rel_l2(dh16.double(), bf(dh_ref.float())) < 3e-3
        assert rel_l2(dx, dz.double() + dh16.double() @ bf(w1)) < 2e-5
        # the ReLU sign as a bit mask (H/8 bytes per row): same dz; mask bits == (h > 0); the dgrad chain driven by the mask,
        # with and without the dh side output, gives the same dx / dh bit for bit as the one driven by the bf16 h
        dz2, h_none, _, _, mask = K.mlp_bwd_ln(x, dout, w1, b1, w2, b2, gamma, want_h=False, want_mask=True)
        assert h_none is None and torch.equal(dz2, dz)
        bits = torch.stack([(mask >> i) & 1 for i in range(64)], dim=2).reshape(R, H).bool()
        assert torch.equal(bits, h16.float() > 0)
        dx2, dh2 = K.mlp_bwd_dgrad(dz, None, w1, w2, mask=mask)
        assert torch.equal(dx2, dx) and torch.equal(dh2, dh16)
        dx3, dh3 = K.mlp_bwd_dgrad(dz, None, w1, w2, mask=mask, want_dh=False)
        assert dh3 is None and torch.equal(dx3, dx)
        # dgrad-only passes skip the dgamma / dbeta column sums: same dz
        dz4, _, g_none, b_none = K.mlp_bwd_ln(x, dout, w1, b1, w2, b2, gamma, want_affine=False)
        assert g_none is None and b_none is None and torch.equal(dz4, dz)


def test_accumulating_stores(cuda_dev):
    """add_ln_bwd(dz_accum=...) and attn_scores_bwd(de_accum=...) add their result into a cotangent that already holds another
    path's contribution (block_backward_backward), instead of a separate elementwise add."""
    R, D, B, N, c = 777, 128, 3, 9, 0.25
    dy, a, b = rnd(cuda_dev, R, D), rnd(cuda_dev, R, D, seed=1), rnd(cuda_dev, R, D, seed=2)
    gamma = rnd(cuda_dev, D, seed=3, scale=0.1) + 1.0
    base = rnd(cuda_dev, R, D, seed=4)
    dz, dgam, dbet = K.add_ln_bwd(dy, a, b, gamma)
    acc = base.clone()
    dz2, dgam2, dbet2 = K.add_ln_bwd(dy, a, b, gamma, dz_accum=acc)
    assert dz2 is acc and torch.equal(acc, base + dz) and rel_l2(dgam2, dgam) < 1e-6 and rel_l2(dbet2, dbet) < 1e-6
    q, k, v = rnd(cuda_dev, B, N, D), rnd(cuda_dev, B, N, D, seed=1), rnd(cuda_dev, B, N, D, seed=2)
    e, dg_, da_in = rnd(cuda_dev, B, N, N, D, seed=3), rnd(cuda_dev, B, N, D, seed=4), rnd(cuda_dev, B, N, N, D, seed=5)
    _, _, stats = K.attn_scores_fwd(q, k, v, e, c, want_stats=True)
    de, dq, dk, dv = K.attn_scores_bwd(dg_, da_in, q, k, v, e, c, stats)
    base = rnd(cuda_dev, B, N, N, D, seed=6)
    acc = base.clone()
    de2, dq2, dk2, dv2 = K.attn_scores_bwd(dg_, da_in, q, k, v, e, c, stats, de_accum=acc)
    assert de2 is acc and torch.equal(acc, base + de) and rel_l2(dq2, dq) < 1e-6 and rel_l2(dk2, dk) < 1e-6
