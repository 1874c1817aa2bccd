"""Raw (non-autograd) kernel entry points of the encoder hot path.

Every function here is one launch (or a fixed short sequence of launches) of a
hand-written sm_100a kernel reached through the C-ABI in ``include/druggen_b200.h``.
Tensors are fp32, contiguous, on one CUDA device; outputs are allocated with the torch
caching allocator and the launch goes onto torch's current stream.  There is NO CPU or
eager fallback: a missing ``libdruggen_b200.so`` or a non-CUDA tensor raises.

Shapes use the reference's vocabulary: B molecules, N atoms (``vertexes``), D = ``dim``
channels; "rows" are flattened node rows [B*N, D] or edge rows [B*N*N, D].
"""
from __future__ import annotations

import os

import torch

from . import _lib

# precision of the dense contractions (GEMMs).  Everything else is always fp32.
#   fp32   : CUDA-core fp32 FMA GEMM (parity mode; bit-for-bit fp32 accumulate)
#   bf16x3 : tcgen05 split precision -- every fp32 operand as hi + lo bf16, three MMAs per product, fp32 accumulation in TMEM
#            (tensor-core parity mode: 16 operand mantissa bits; everything stays fp32 in HBM)
#   bf16   : tcgen05 bf16 x bf16 -> fp32 (TMEM accumulators), fused chains, bf16 storage of operand-only tensors -- throughput mode
PRECISIONS = ("fp32", "bf16x3", "bf16")
# The reference trains in fp32, so the library default is the tensor-core PARITY mode (gradients at the reference's own fp32
# noise level, tests/test_parity_gpu.py::test_gan_step_depth8_n45_vs_oracle); the bf16 throughput mode is an explicit opt-in
# (bench.py --precision bf16, DRUGGEN_B200_PRECISION=bf16, ``with druggen_b200.precision("bf16")``).
_precision = os.environ.get("DRUGGEN_B200_PRECISION", "bf16x3")


def set_precision(p: str) -> None:
    global _precision
    if p not in PRECISIONS:
        raise ValueError(f"precision must be one of {PRECISIONS}, got {p!r}")
    _precision = p


def get_precision() -> str:
    return _precision


class precision:
    """Context manager: ``with kernels.precision("fp32"): ...``"""

    def __init__(self, p):
        self.p = p

    def __enter__(self):
        self.old = get_precision()
        set_precision(self.p)

    def __exit__(self, *a):
        set_precision(self.old)


_test_backend = None
_debug_hold = []


def _install_backend_for_tests(backend) -> None:
    """tests/ only: swap the launch table for a torch emulation so the autograd wiring can be
    checked on a box without a GPU.  Never called by product code."""
    global _test_backend
    _test_backend = backend


def _be():
    if _test_backend is not None:
        return _test_backend
    return _lib.cuda_backend()


def _chk(*ts, bf16_ok=False):
    if _test_backend is not None:
        return
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("druggen_b200 kernels run on CUDA tensors only (no CPU fallback)")
        if dev is None:
            dev = t.device
            _lib.cuda_backend().device = dev       # the launch goes to the tensors' device, not merely the current one
        elif t.device != dev:
            raise RuntimeError(f"druggen_b200 kernels take tensors of one device, got {dev} and {t.device}")
        if t.dtype != torch.float32 and not (bf16_ok and t.dtype == torch.bfloat16):
            raise RuntimeError(f"druggen_b200 kernels take fp32 tensors, got {t.dtype}")
        if not t.is_contiguous():
            raise RuntimeError("druggen_b200 kernels take contiguous tensors")


# ----------------------------------------------------------------------------- dense contractions
def rows_gemm(a, w, w_is_nk: bool, bias=None, relu: bool = False, gate=None, resid=None, out_bf16: bool = False):
    """out[R,N] = epi(a[R,K] . op(w) + bias) + resid.

    w_is_nk: w is [N,K] (nn.Linear layout, out = a w^T) else w is [K,N] (out = a w).
    relu: clamp at 0.  gate: optional [R,N]; out *= (gate > 0)   (ReLU backward fused).
    resid: optional [R,N] added in the store (gradient accumulation without an extra pass).
    bf16 storage (tensor-core mode, shapes the tcgen05 kernel takes): ``a`` and ``gate`` may be bf16 tensors and
    ``out_bf16`` stores the result as bf16 -- for tensors that are only ever contraction operands or sign masks
    (the MLP hidden activation and its gradient) this loses nothing: the contraction rounds them to bf16 anyway.
    """
    _chk(w, bias, resid)
    _chk(a, gate, bf16_ok=True)
    r, k = a.shape
    n = w.shape[0] if w_is_nk else w.shape[1]
    assert (w.shape[1] if w_is_nk else w.shape[0]) == k, (a.shape, w.shape, w_is_nk)
    out = torch.empty((r, n), dtype=torch.bfloat16 if out_bf16 else w.dtype, device=a.device)
    if r:
        _be().rows_gemm(a, w, w_is_nk, bias, relu, gate, out, _precision, resid)
    return out


def gemm_tn(a, b, out=None, colsum_a=None):
    """out[M,N] (+)= a[R,M]^T . b[R,N]  -- weight-gradient contraction over rows.
    ``out`` given => accumulate into it.  ``colsum_a`` [M] given => += column sums of a (bias gradient,
    same pass over a)."""
    _chk(out, colsum_a)
    _chk(a, b, bf16_ok=True)
    assert a.shape[0] == b.shape[0]
    acc = out is not None
    if out is None:
        wide = [t.dtype for t in (a, b) if t.dtype != torch.bfloat16]
        out = torch.zeros((a.shape[1], b.shape[1]), dtype=wide[0] if wide else torch.float32, device=a.device)
    if a.shape[0]:
        _be().gemm_tn(a, b, out, acc, _precision, colsum_a)
    return out


def colsum(a):
    """out[N] = sum over rows of a[R,N] (bias gradient)."""
    _chk(a)
    out = a.new_zeros((a.shape[1],))
    if a.shape[0]:
        _be().colsum(a, out)
    return out


def gate_mul(x, ref):
    """x * (ref > 0)."""
    _chk(x, ref)
    out = torch.empty_like(x)
    if x.numel():
        _be().gate_mul(x, ref, out)
    return out


# ----------------------------------------------------------------------------- residual + LayerNorm
def add_ln_fwd(a, b, gamma, beta, eps: float = 1e-5):
    """LN(a + b) * gamma + beta over the last dim; b may be None."""
    _chk(a, b, gamma, beta)
    out = torch.empty_like(a)
    if a.numel():
        _be().add_ln_fwd(a, b, gamma, beta, out, eps)
    return out


def add_ln_bwd(dy, a, b, gamma, eps: float = 1e-5, dz_accum=None):
    """-> (dz, dgamma, dbeta) with z = a + b recomputed.  ``dz_accum`` given => dz is added into it in place (and returned)."""
    _chk(dy, a, b, gamma, dz_accum)
    dz = torch.empty_like(a) if dz_accum is None else dz_accum
    dgamma = torch.zeros_like(gamma)
    dbeta = torch.zeros_like(gamma)
    if a.numel():
        _be().add_ln_bwd(dy, a, b, gamma, dz, dgamma, dbeta, eps, accumulate=dz_accum is not None)
    return dz, dgamma, dbeta


def add_ln_bwd_bwd(u, vg, vb, dy, a, b, gamma, eps: float = 1e-5):
    """Gradient of <u,dz> + <vg,dgamma> + <vb,dbeta> w.r.t. (dy, z, gamma).  vg/vb may be None."""
    _chk(u, vg, vb, dy, a, b, gamma)
    g_dy = torch.empty_like(a)
    g_z = torch.empty_like(a)
    g_gamma = torch.zeros_like(gamma)
    if a.numel():
        _be().add_ln_bwd_bwd(u, vg, vb, dy, a, b, gamma, g_dy, g_z, g_gamma, eps)
    return g_dy, g_z, g_gamma


# ----------------------------------------------------------------------------- edge-modulated scores
def modulate_fwd(q, k, e, c: float):
    """A[b,i,j,:] = c * q[b,i,:] * k[b,j,:] * (e^2 + e)[b,i,j,:]   (layers.py:123-125)."""
    _chk(q, k, e)
    out = torch.empty_like(e)
    if e.numel():
        _be().modulate_fwd(q, k, e, c, out)
    return out


def modulate_bwd(da, q, k, e, c: float):
    """-> (dq, dk, de)."""
    _chk(da, q, k, e)
    dq, dk, de = torch.empty_like(q), torch.zeros_like(k), torch.empty_like(e)
    if e.numel():
        _be().modulate_bwd(da, q, k, e, c, dq, dk, de)
    return dq, dk, de


def modulate_bwd_bwd(uq, uk, ue, da, q, k, e, c: float):
    """Gradient of <uq,dq>+<uk,dk>+<ue,de> w.r.t. (da, q, k, e)."""
    _chk(uq, uk, ue, da, q, k, e)
    g_da, g_e = torch.empty_like(e), torch.empty_like(e)
    g_q, g_k = torch.empty_like(q), torch.zeros_like(k)
    if e.numel():
        _be().modulate_bwd_bwd(uq, uk, ue, da, q, k, e, c, g_da, g_q, g_k, g_e)
    return g_da, g_q, g_k, g_e


# ----------------------------------------------------------------------------- softmax over keys + aggregate
def softmax_agg_fwd(a, v):
    """g[b,i,:] = sum_j softmax_j(a[b,i,j,:]) * v[b,j,:]   (layers.py:130-134)."""
    _chk(a, v)
    out = torch.empty_like(v)
    if a.numel():
        _be().softmax_agg_fwd(a, v, out)
    return out


def softmax_agg_bwd(dg, a, v, da_accum=None):
    """-> (da, dv).  ``da_accum`` given => the result is added into it in place (and returned)."""
    _chk(dg, a, v, da_accum)
    da = torch.empty_like(a) if da_accum is None else da_accum
    dv = torch.zeros_like(v)
    if a.numel():
        _be().softmax_agg_bwd(dg, a, v, da, dv, da_accum is not None)
    return da, dv


def softmax_agg_bwd_bwd(ua, uv, dg, a, v):
    """Gradient of <ua,da>+<uv,dv> w.r.t. (dg, a, v)."""
    _chk(ua, uv, dg, a, v)
    g_dg, g_a, g_v = torch.empty_like(dg), torch.empty_like(a), torch.zeros_like(v)
    if a.numel():
        _be().softmax_agg_bwd_bwd(ua, uv, dg, a, v, g_dg, g_a, g_v)
    return g_dg, g_a, g_v


# ----------------------------------------------------------------------------- fused tcgen05 kernels
def fused_available(d: int, h: int) -> bool:
    """The fused tensor-core kernels exist for the throughput mode and the reference's width."""
    return _precision == "bf16" and d == 128 and h % 128 == 0 and 128 <= h <= 384


def mlp_fwd(x, w1, b1, w2, b2, gamma, beta, eps: float = 1e-5):
    """LN(x + fc2(relu(fc1(x)+b1)) + b2) * gamma + beta in one kernel; x:[R,128]."""
    _chk(x, w1, b1, w2, b2, gamma, beta)
    out = torch.empty_like(x)
    if x.numel():
        ws = torch.empty(2 * (w1.shape[0] // 128) * 32768, dtype=torch.uint8, device=x.device)
        _be().mlp_fwd(x, w1, b1, w2, b2, gamma, beta, out, eps, ws)
    return out


def _mlp_ws(w1):
    return torch.empty(2 * (w1.shape[0] // 128) * 32768, dtype=torch.uint8, device=w1.device)


def mlp_bwd_ln(x, dout, w1, b1, w2, b2, gamma, eps: float = 1e-5, want_h: bool = True, want_mask: bool = False,
               want_affine: bool = True):
    """Recompute the residual MLP from x and take the LayerNorm backward of ``dout``.
    -> (dz [R,128] fp32, h [R,H] bf16 | None, dgamma, dbeta[, mask]).  ``want_mask``: also the ReLU sign mask [R, H/64] int64
    (H/8 bytes per row) -- all ``mlp_bwd_dgrad`` needs of h; ``want_h=False`` skips the bf16 h (only the fc2 weight gradient
    reads it)."""
    _chk(x, dout, w1, b1, w2, b2, gamma)
    dz = torch.empty_like(x)
    h16 = torch.empty((x.shape[0], w1.shape[0]), dtype=torch.bfloat16, device=x.device) if want_h else None
    mask = torch.empty((x.shape[0], w1.shape[0] // 64), dtype=torch.int64, device=x.device) if want_mask else None
    dgamma, dbeta = (torch.zeros_like(gamma), torch.zeros_like(gamma)) if want_affine else (None, None)   # (None: column sums skipped)
    if x.numel():
        _be().mlp_bwd_ln(x, dout, w1, b1, w2, b2, gamma, dz, h16, dgamma, dbeta, eps, _mlp_ws(w1), mask=mask)
    return (dz, h16, dgamma, dbeta, mask) if want_mask else (dz, h16, dgamma, dbeta)


def mlp_bwd_dgrad(dz, h16, w1, w2, mask=None, want_dh: bool = True):
    """-> (dx = dz + dh.W1 [R,128] fp32, dh = (dz.W2)*(h>0) [R,H] bf16 | None).  The sign of h comes from ``mask``
    (``mlp_bwd_ln(want_mask=True)``) when given, else from ``h16``; ``want_dh=False`` skips the bf16 dh (only the fc1
    weight gradient reads it)."""
    _chk(dz, w1, w2)
    _chk(h16, bf16_ok=True)
    dx = torch.empty_like(dz)
    dh16 = torch.empty((dz.shape[0], w1.shape[0]), dtype=torch.bfloat16, device=dz.device) if want_dh else None
    if dz.numel():
        _be().mlp_bwd_dgrad(dz, h16, w1, w2, dx, dh16, _mlp_ws(w1), mask=mask)
    return dx, dh16


# ----------------------------------------------------------------------------- fused attention scores (fp32)
def attn_fused_available(n: int, d: int, b: int = 1) -> bool:
    """Shapes the fused per-molecule score kernels take (the host-side limits of ``attn_ok`` in csrc/attn_scores.cu): D = 128,
    N >= 4, the per-CTA dk / dv accumulators (2 N + 16) * 512 B within 220 KB (N <= 212), B within the grid.y limit.
    Anything else runs the generic modulate / softmax-aggregate kernels (still CUDA, never a CPU path)."""
    return d == 128 and 4 <= n <= 212 and 0 < b <= 65535


def attn_scores_fwd(q, k, v, e, c: float, want_stats: bool = False, store_a: bool = True):
    """-> (a, g[, stats]): modulated scores (layers.py:123-125) and softmax-aggregate (layers.py:130-134), e read once.
    ``want_stats`` also returns (max, 1/sum, g) per (b, i, channel) for ``attn_scores_bwd``.
    ``store_a=False``: the scores are not written (a is None) -- only g and the statistics are wanted."""
    _chk(q, k, v, e)
    a, g = (torch.empty_like(e) if store_a else None), torch.empty_like(q)
    stats = (torch.empty_like(q), torch.empty_like(q)) if want_stats else None
    if e.numel():
        _be().attn_scores_fwd(q, k, v, e, c, a, g, stats)
    return (a, g, stats + (g,)) if want_stats else (a, g)


def attn_scores_bwd(dg, da_in, q, k, v, e, c: float, stats=None, de_bf16: bool = False, scores_bf16: bool = False, de_accum=None):
    """-> (de, dq, dk, dv) from dg (softmax path) and da_in (out_e path; may be None).  ``stats`` from the forward
    skips the statistics sweep.  ``de_bf16``: de is stored as bf16 (it is only ever a contraction operand).
    ``scores_bf16``: the statistics were taken from bf16-stored scores (softmax_agg16_fwd); the recomputed scores are rounded alike."""
    _chk(dg, q, k, v, e)
    _chk(da_in, bf16_ok=True)          # (bf16 da: the out_e dgrad GEMM stored it with out_bf16 -- tensor-core mode)
    _chk(de_accum)                     # (fp32, same shape as e: de is added into it in place and returned)
    de = de_accum if de_accum is not None else torch.empty_like(e, dtype=torch.bfloat16 if de_bf16 else e.dtype)
    dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
    if e.numel():
        _be().attn_scores_bwd(dg, da_in, q, k, v, e, c, de, dq, dk, dv, stats, scores_bf16, accumulate_de=de_accum is not None)
    return de, dq, dk, dv


# ----------------------------------------------------------------------------- fused tcgen05 edge-attention chain
def attn_chain_available(b: int, n: int, d: int) -> bool:
    """E-projection -> modulation -> out_e projection -> residual -> LN4 as one tcgen05 kernel (throughput mode)."""
    # (host-side limits of dg_attn_edge_fwd / dg_softmax_agg16_fwd: 32-bit row and element offsets, grid.y)
    return (_precision == "bf16" and d == 128 and 4 <= n <= 212 and 0 < b <= 65535 and b * n * n < 2 ** 31 and b * n * 128 < 2 ** 31
            and os.environ.get("DRUGGEN_B200_ATTN_CHAIN", "1") != "0")


def softmax_scores_bf16() -> bool:
    """Throughput-mode choice for the softmax over key atoms on the fused-chain path: True (default) = its input is the
    bf16 copy of the scores that the chain kernel spills anyway (256 B per edge row; the exp amplifies the 2^-9 relative
    rounding to |a| * 2^-9 relative error in each probability); DRUGGEN_B200_SOFTMAX_SCORES=fp32 = the chain spills
    E in fp32 and the softmax recomputes the scores in fp32 (+~5 % step time), as the non-chain path always does."""
    return os.environ.get("DRUGGEN_B200_SOFTMAX_SCORES", "bf16") != "fp32"


def attn_edge_fwd(y, q, k, we, be, woe, boe, gamma, beta, c: float, want_a16: bool = True, want_e: bool = False,
                  want_z: bool = False, eps: float = 1e-5):
    """y:[B*N*N,128], q,k:[B,N,128] -> (y3 = LN4(y + out_e(A)), a16 | None, e | None, z | None)   (layers.py:116,123-127,188,190).
    a16: the scores A = c q_i k_j (E^2+E) as bf16; e: E = y We^T + be fp32; z: y + out_e(A) fp32 (input of LN4)."""
    _chk(y, q, k, we, be, woe, boe, gamma, beta)
    b, n, d = q.shape
    assert y.shape == (b * n * n, d), (y.shape, q.shape)
    if os.environ.get("DG_DEBUG_ATTN_UNFUSED"):          # debug: the same outputs from the unfused kernels
        e_ = rows_gemm(y, we, True, be)
        a_ = modulate_fwd(q, k, e_.view(b, n, n, d), c).view(-1, d)
        y1 = rows_gemm(a_, woe, True, boe)
        return (add_ln_fwd(y, y1, gamma, beta), a_.to(torch.bfloat16) if want_a16 else None, e_ if want_e else None,
                (y + y1) if want_z else None)
    out = torch.empty_like(y)
    a16 = torch.empty_like(y, dtype=torch.bfloat16) if want_a16 else None
    e = torch.empty_like(y) if want_e else None
    z = torch.empty_like(y) if want_z else None
    if y.numel():
        ws = torch.empty(2 * 32768, dtype=torch.uint8, device=y.device)
        _be().attn_edge_fwd(y, q, k, we, be, woe, boe, gamma, beta, c, out, a16, e, z, eps, ws)
    if os.environ.get("DG_DEBUG_HOLD"):                    # debug: keep every output alive (no allocator reuse of their storage)
        _debug_hold.append((out, a16, e, z, ws if y.numel() else None)[: int(os.environ["DG_DEBUG_HOLD"])])
        _debug_hold.append(("clone", out, out.clone(), y, y.clone()))
    if os.environ.get("DG_DEBUG_ATTN_COMPARE"):           # debug: compare every output with the unfused kernels
        e_ = rows_gemm(y, we, True, be)
        a_ = modulate_fwd(q, k, e_.view(b, n, n, d), c).view(-1, d)
        y1 = rows_gemm(a_, woe, True, boe)
        rl = lambda g, r: float((g.float() - r.float()).norm() / r.float().norm().clamp_min(1e-30))  # noqa: E731
        msg = "attn_edge_fwd[%s%s%s] y3 %.2e" % ("a" if want_a16 else "", "e" if want_e else "", "z" if want_z else "",
                                                  rl(out, add_ln_fwd(y, y1, gamma, beta)))
        ref_out = add_ln_fwd(y, y1, gamma, beta)
        dd = (out - ref_out).abs()
        msg += " maxabs %.2e n>1e-5 %d argmax_row %d col %d" % (float(dd.max()), int((dd > 1e-5).sum()), int(dd.max(1).values.argmax()),
                                                                 int(dd.max(0).values.argmax()))
        if want_a16:
            msg += " a16 %.2e" % rl(a16, a_.to(torch.bfloat16))
        if want_e:
            msg += " E %.2e" % rl(e, e_)
        if want_z:
            msg += " z %.2e" % rl(z, y + y1)
        if os.environ.get("DG_DEBUG_ATTN_QUIET") is None:
            print(msg + "  |y| %.2e nonfinite %d" % (float(y.abs().max()), int((~torch.isfinite(out)).sum())))
        sub = int(os.environ.get("DG_DEBUG_ATTN_SUBST", "0"))
        if sub & 1:
            out = add_ln_fwd(y, y1, gamma, beta)
            if os.environ.get("DG_DEBUG_NOISE"):
                out = out * (1.0 + float(os.environ["DG_DEBUG_NOISE"]) * torch.randn_like(out))
        if sub & 2 and want_a16:
            a16 = a_.to(torch.bfloat16)
        if sub & 4 and want_e:
            e = e_
        if sub & 8 and want_z:
            z = y + y1
    return out, a16, e, z


def softmax_agg16_fwd(a16, v, want_stats: bool = False):
    """g[b,i,:] = sum_j softmax_j(a16[b,i,j,:]) v[b,j,:] from bf16 scores (fp32 arithmetic) -> g | (g, (max, 1/sum, g))."""
    _chk(v)
    _chk(a16, bf16_ok=True)
    g = torch.empty_like(v)
    stats = (torch.empty_like(v), torch.empty_like(v)) if want_stats else None
    if a16.numel():
        _be().softmax_agg16_fwd(a16, v, g, stats)
    return (g, stats + (g,)) if want_stats else g


# ----------------------------------------------------------------------------- block-level entry points
def native_block_available(b: int, n: int, d: int, h: int) -> bool:
    """One C call per direction of an encoder block (``dg_block_fwd`` / ``dg_block_bwd``: the library sequences its own launches
    over buffers allocated here) -- exists for the throughput mode on the fused-chain path; everything else, and bench.py's
    per-launch kernel table, runs the same launches one by one from ``block.py``.  DRUGGEN_B200_NATIVE_BLOCK=0 switches it off."""
    return (_test_backend is None and fused_available(d, h) and attn_chain_available(b, n, d) and softmax_scores_bf16()
            and os.environ.get("DRUGGEN_B200_NATIVE_BLOCK", "1") != "0" and not os.environ.get("DG_DEBUG_ATTN_UNFUSED")
            and not os.environ.get("DG_DEBUG_ATTN_COMPARE") and not os.environ.get("DG_DEBUG_HOLD") and _lib.cuda_backend().native_blocks())


def _chk_buffers(ts, dev):
    for t in ts:
        if t is not None and (t.device != dev or not t.is_contiguous()):
            raise RuntimeError("druggen_b200 block entry points take contiguous tensors of the parameters' device")


def block_fwd(io: dict, params, b: int, n: int, d: int, h: int, heads: int, flags: int, eps: float = 1e-5) -> None:
    """``dg_block_fwd`` over the named buffers of ``io`` (``_lib.BLK_SLOTS``)."""
    _chk(*params)
    _chk_buffers(io.values(), params[0].device)
    _be().block_fwd(io, params, b, n, d, h, heads, flags, eps, _mlp_ws(params[18]))


def block_bwd(io: dict, params, grads, b: int, n: int, d: int, h: int, heads: int, flags: int, eps: float = 1e-5) -> None:
    """``dg_block_bwd`` over the named buffers of ``io``; ``grads``: 30 zeroed tensors (None where no consumer) or None."""
    _chk(*params)                 # (the gradient tensors are the caller's own fresh views: block._flat_grads)
    _chk_buffers(io.values(), params[0].device)
    _be().block_bwd(io, params, grads, b, n, d, h, heads, flags, eps, _mlp_ws(params[18]))


def block_bwd_bwd(io: dict, params, grads, b: int, n: int, d: int, h: int, heads: int, flags: int, eps: float = 1e-5) -> None:
    """``dg_block_bwd_bwd`` (the gradient penalty's second-order pass of one block) over the named buffers of ``io``."""
    _chk(*params)
    _chk_buffers(io.values(), params[0].device)
    _be().block_bwd_bwd(io, params, grads, b, n, d, h, heads, flags, eps, _mlp_ws(params[18]))


def encoder_fwd(x, y, x_out, y_out, params, depth: int, scratch: dict, b: int, n: int, d: int, h: int, heads: int, last_edge_out: bool,
                eps: float = 1e-5, ws=None) -> None:
    """``dg_encoder_fwd``: ``depth`` blocks in one call; ``params`` = depth x 30 tensors, ``scratch`` = named DG_BLK_* buffers;
    ``ws``: the packed-weight workspace (``_mlp_ws``), allocated here unless given (CUDA-graph capture: nothing may be allocated)."""
    _chk(x, y, x_out, y_out, *params)
    _chk_buffers(scratch.values(), x.device)
    _be().encoder_fwd(x, y, x_out, y_out, params, depth, scratch, b, n, d, h, heads, last_edge_out, eps,
                      _mlp_ws(params[18]) if ws is None else ws)


def set_option(key: int, value: int) -> None:
    """Process-wide tuning switches of the library (``_lib.OPT_*``)."""
    if _test_backend is None:
        _lib.load().dg_set_option(key, value)


# ----------------------------------------------------------------------------- either side of the encoder path
def _chk_labels(labels):
    if _test_backend is None and not labels.is_cuda:
        raise RuntimeError("druggen_b200 kernels run on CUDA tensors only (no CPU fallback)")
    if labels.dtype not in (torch.int64, torch.uint8):
        raise RuntimeError(f"labels are int64 or uint8, got {labels.dtype}")
    return labels.contiguous()


def check_labels() -> None:
    """Raise if any label-consuming launch since the last check met a label outside [0, classes) -- where the reference's
    ``scatter_`` (src/data/utils.py:21) raises.  Synchronises with the device."""
    flag = _be().label_error(True) if _test_backend is None else 0
    if flag > 0 and flag & 2:
        raise RuntimeError("druggen_b200: node or graph index out of range in to_dense_adj (edge_index / batch)")
    if flag > 0:
        raise RuntimeError("druggen_b200: label outside [0, classes) (the reference's label2onehot scatter_ raises here)")


def label2onehot(labels, dim: int, device=None, validate: bool = True):
    """src/data/utils.py:15-23, same signature: integer labels [...] -> fp32 one-hot [..., dim] on the labels' CUDA device
    (``device`` given: labels are moved there first -- as uint8 this is the 1-byte-per-edge wire format).  Like the
    reference's ``scatter_`` an out-of-range label raises (``validate=False`` defers that to ``check_labels()``: no sync here)."""
    if device is not None:
        labels = labels.to(device, non_blocking=True)
    labels = _chk_labels(labels)
    out = torch.empty(tuple(labels.shape) + (dim,), dtype=torch.float32, device=labels.device)
    if labels.numel():
        _be().label2onehot(labels, out, dim)
        if validate:
            check_labels()
    return out


def to_dense_adj(edge_index, batch, edge_attr=None, max_num_nodes: int = None, batch_size: int = None):
    """torch_geometric.utils.to_dense_adj (PyG 2.2.0) for integer edge attributes, as load_molecules calls it
    (src/data/utils.py:130-135): -> int32 [B, N, N].  ``batch_size`` / ``max_num_nodes`` default to PyG's (batch.max() + 1 and
    the largest graph: both need a device sync; load_molecules passes them)."""
    _chk_ints(edge_index, batch, edge_attr)
    if batch_size is None:
        batch_size = int(batch.max()) + 1 if batch.numel() else 1
    if max_num_nodes is None:
        max_num_nodes = int(torch.bincount(batch, minlength=batch_size).max()) if batch.numel() else 0
    adj = torch.empty(batch_size, max_num_nodes, max_num_nodes, dtype=torch.int32, device=batch.device)
    cum = torch.empty(batch_size + 1, dtype=torch.int64, device=batch.device)
    if adj.numel():
        _be().to_dense_adj(edge_index.contiguous(), batch.contiguous(), None if edge_attr is None else edge_attr.contiguous(), adj, cum)
    return adj


def _chk_ints(*ts):
    for t in ts:
        if t is None:
            continue
        if _test_backend is None and not t.is_cuda:
            raise RuntimeError("druggen_b200 kernels run on CUDA tensors only (no CPU fallback)")
        if t.dtype != torch.int64:
            raise RuntimeError(f"edge_index / batch / edge_attr are int64 (PyG's layout), got {t.dtype}")


def narrow_labels(adj, classes: int, validate: bool = True):
    """int32 labels -> uint8 (the 1-byte wire format); values outside [0, classes) raise like label2onehot's scatter_."""
    if _test_backend is None and not adj.is_cuda:
        raise RuntimeError("druggen_b200 kernels run on CUDA tensors only (no CPU fallback)")
    assert adj.dtype == torch.int32 and adj.is_contiguous()
    out = torch.empty(adj.shape, dtype=torch.uint8, device=adj.device)
    if adj.numel():
        _be().narrow_labels(adj, out, classes)
        if validate:
            check_labels()
    return out


def pack_bits(vecs):
    """0/1 fingerprint matrix [rows, F] (uint8 or float32; non-zero = set) -> (bits uint64-as-int64 [rows, W], popcounts int32 [rows]),
    W = F / 64 rounded up to a power of two (zero words appended)."""
    if _test_backend is None and not vecs.is_cuda:
        raise RuntimeError("druggen_b200 kernels run on CUDA tensors only (no CPU fallback)")
    assert vecs.dim() == 2 and vecs.dtype in (torch.uint8, torch.float32) and vecs.is_contiguous(), (vecs.shape, vecs.dtype)
    rows, f = vecs.shape
    w0 = (f + 63) // 64
    w = 1
    while w < w0:
        w *= 2
    if w > 32:
        raise RuntimeError(f"fingerprints of {f} bits exceed the kernel's 2048")
    bits = torch.zeros(rows, w, dtype=torch.int64, device=vecs.device)
    cnt = torch.zeros(rows, dtype=torch.int32, device=vecs.device)
    if rows:
        if w == w0:
            _be().pack_bits(vecs, bits, cnt)
        else:                                   # (packed at the natural width, then copied into the zero-padded rows)
            tight = torch.zeros(rows, w0, dtype=torch.int64, device=vecs.device)
            _be().pack_bits(vecs, tight, cnt)
            bits[:, :w0] = tight
    return bits, cnt


def tanimoto_agg(stock, gen, agg: str = "max", p: float = 1.0):
    """stock / gen = ``pack_bits`` results.  agg 'max' -> float32 [G] max_s jac(s, g) (>= 0);  agg 'sum' -> float64 [G] sum_s jac^p."""
    (sb, sc), (gb, gc) = stock, gen
    assert sb.shape[1] == gb.shape[1], "fingerprint widths differ"
    g = gb.shape[0]
    out_max = torch.zeros(g, dtype=torch.float32, device=gb.device) if agg == "max" else None
    out_sum = torch.zeros(g, dtype=torch.float64, device=gb.device) if agg != "max" else None
    if g and sb.shape[0]:
        _be().tanimoto_agg(sb, sc, gb, gc, 0 if agg == "max" else 1, float(p), out_max, out_sum)
    return out_max if agg == "max" else out_sum


def argmax_last(t):
    """inference.py:197-198 ``torch.max(t, -1)[1]``: int64 index of the first maximum over the last dim."""
    _chk(t)
    out = torch.empty(t.shape[:-1], dtype=torch.int64, device=t.device)
    if t.numel():
        _be().argmax_last(t, out)
    return out


def symmetrize(e):
    """models.py:94: (e + e.permute(0, 2, 1, 3)) / 2 for e:[B,N,N,D] in one pass."""
    _chk(e)
    assert e.dim() == 4 and e.shape[1] == e.shape[2], e.shape
    out = torch.empty_like(e)
    if e.numel():
        _be().symmetrize(e, out)
    return out


def embed_labels_fwd(labels, lut, sym: bool):
    """Prologue of a one-hot batch given as labels (models.py:91-94): labels [B,N] -> lut[labels] [B,N,D]; ``sym`` (edges,
    labels [B,N,N]) -> (lut[a_ij] + lut[a_ji]) / 2 [B,N,N,D].  lut:[classes,D] = the prologue MLP applied to the identity."""
    _chk(lut)
    labels = _chk_labels(labels)
    assert labels.dim() == (3 if sym else 2) and (not sym or labels.shape[1] == labels.shape[2]), labels.shape
    y = torch.empty(tuple(labels.shape) + (lut.shape[1],), dtype=lut.dtype, device=lut.device)
    if labels.numel():
        _be().embed_labels_fwd(labels, lut, y, labels.shape[1], sym)
    return y


def embed_labels_bwd(labels, dy, classes: int, sym: bool):
    """-> dlut [classes,D]: segmented sum of dy rows by label (both labels of the symmetrised pair get half)."""
    _chk(dy)
    labels = _chk_labels(labels)
    dlut = torch.zeros((classes, dy.shape[-1]), dtype=dy.dtype, device=dy.device)
    if labels.numel():
        _be().embed_labels_bwd(labels, dy, dlut, labels.shape[1], sym)
    return dlut


def gp_interp(labels, fake, eps):
    """loss.py:21-26 ``eps * real + (1 - eps) * fake`` with ``real`` as labels [B,...] and ``fake`` [B,...,classes];
    eps:[B] (or [B,1,..]).  Bit-identical to torch's elementwise kernels on the one-hot tensor."""
    _chk(fake, eps)
    labels = _chk_labels(labels)
    assert tuple(fake.shape[:-1]) == tuple(labels.shape), (fake.shape, labels.shape)
    out = torch.empty_like(fake)
    if labels.numel():
        _be().gp_interp(labels, fake, eps.reshape(-1), out, labels.numel() // labels.shape[0])
    return out


def gp_penalty(g_node, g_edge):
    """loss.py:42-47 -> (penalty [1], coef [B]): mean_b (|concat(g_node_b, g_edge_b)| - 1)^2 and d penalty / d g = coef_b g."""
    _chk(g_node, g_edge)
    b = g_node.shape[0]
    pen = torch.empty(1, dtype=g_node.dtype, device=g_node.device)
    coef = torch.empty(b, dtype=g_node.dtype, device=g_node.device)
    _be().gp_penalty(g_node, g_edge, pen, coef, torch.empty_like(coef))
    return pen, coef


def gp_penalty_bwd(g, coef, upstream):
    _chk(g, coef, upstream)
    out = torch.empty_like(g)
    _be().gp_penalty_bwd(g, coef, upstream.reshape(1), out)
    return out


def readout_argmax(x, w, bias, want_logits: bool = False, idx_dtype=torch.int64):
    """models.py:100-101 + inference.py:197-198: x:[...,128] -> (idx [...], logits [...,classes] | None) in one pass:
    logits = x w^T + bias, idx = their first maximum (int64, or uint8 for a 1-byte-per-edge result)."""
    _chk(x, w, bias)
    assert idx_dtype in (torch.int64, torch.uint8)
    idx = torch.empty(x.shape[:-1], dtype=idx_dtype, device=x.device)
    logits = torch.empty(tuple(x.shape[:-1]) + (w.shape[0],), dtype=x.dtype, device=x.device) if want_logits else None
    if x.numel():
        _be().readout_argmax(x, w, bias, logits, idx)
    return idx, logits


def adamw_flat(p, g, m, v, segs, nseg: int, lr: float, beta1: float, beta2: float, eps: float, weight_decay: float) -> None:
    """One fused AdamW launch over a network's flat parameter / gradient / moment buffers (train.py:213-214)."""
    _chk(p, g, m, v)
    _be().adamw_flat(p, g, m, v, segs, nseg, lr, beta1, beta2, eps, weight_decay)
