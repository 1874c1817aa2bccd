#!/usr/bin/env python
"""Micro-benchmarks of individual hot-path kernels on one GPU (CUDA events, L2-exceeding inputs).

    python tools/kernel_bench.py [--batch 512] [--atoms 45]

Prints one JSON line per kernel: ms, achieved GB/s (algorithmic bytes) and TFLOP/s, and the
fraction of the measured peaks in MEASURED_PEAKS.json.  Also the encoder-only forward
(BASELINE config 5) as molecules/s.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import druggen_b200 as dg  # noqa: E402
from druggen_b200 import kernels as K  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--atoms", type=int, default=45)
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--keep-ab", action="store_true", help="time every kernel without / with L2::evict_last on the chain loaders' x reads")
    ap.add_argument("--pf", type=int, default=None, help="A/B: the default L2-prefetch mask against this one")
    ap.add_argument("--prefetch-ab", action="store_true", help="time every kernel with the L2-prefetch option off and on")
    args = ap.parse_args()
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pk = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    dev = torch.device("cuda:0")
    b, n, d, h = args.batch, args.atoms, 128, 384
    r = b * n * n
    g = torch.Generator(device="cpu").manual_seed(0)
    rn = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)  # noqa: E731
    x = rn(r, d)
    w1, b1, w2, b2 = rn(h, d, sc=d ** -0.5), rn(h, sc=0.1), rn(d, h, sc=h ** -0.5), rn(d, sc=0.1)
    gamma, beta = torch.ones(d, device=dev), torch.zeros(d, device=dev)

    def report(name, ms, nbytes, flops, **extra):
        rec = {"kernel": name, "rows": r, "ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1),
               "hbm_frac": round(nbytes / ms / 1e6 / pk["hbm_gbs"], 3), "TFLOPs": round(flops / ms / 1e9, 1),
               "tensor_frac": round(flops / ms / 1e9 / pk["bf16_tflops"], 3)}
        rec.update(extra)
        print(json.dumps(rec), flush=True)

    from druggen_b200 import _lib
    dout = rn(r, d)
    w = rn(d, d, sc=d ** -0.5)
    q, k, v = rn(b, n, d), rn(b, n, d), rn(b, n, d)
    dgn = rn(b, n, d)
    acc4 = torch.zeros(b, n, n, d, device=dev)
    x4 = x.view(b, n, n, d)

    with dg.precision("bf16"):
        dz, h16, _, _, mask = K.mlp_bwd_ln(x, dout, w1, b1, w2, b2, gamma, want_mask=True)
        _, a16, _, _ = K.attn_edge_fwd(x, q, k, w, b2, w, b2, gamma, beta, 0.25)
        _, _, stats = K.attn_scores_fwd(q, k, v, x4, 0.25, want_stats=True, store_a=False)

        def unfused():
            hh = K.rows_gemm(x, w1, True, b1, True)
            m = K.rows_gemm(hh, w2, True, b2)
            return K.add_ln_fwd(x, m, gamma, beta)

        def unfused_attn_edge():
            e = K.rows_gemm(x, w, True, b2)
            a, g = K.attn_scores_fwd(q, k, v, e.view(b, n, n, d), 0.25)
            y1 = K.rows_gemm(a.view(-1, d), w, True, b2)
            return K.add_ln_fwd(x, y1, gamma, beta)

        def fused_attn_edge():
            y3, a16_, _, _ = K.attn_edge_fwd(x, q, k, w, b2, w, b2, gamma, beta, 0.25)
            return K.softmax_agg16_fwd(a16_, v)

        benches = [
            ("mlp_fwd[fused,H=384]", lambda: K.mlp_fwd(x, w1, b1, w2, b2, gamma, beta), 2 * r * d * 4, 4.0 * r * d * h),
            ("mlp_bwd_ln[fused,H=384]", lambda: K.mlp_bwd_ln(x, dout, w1, b1, w2, b2, gamma), r * (3 * d * 4 + h * 2), 4.0 * r * d * h),
            ("mlp_bwd_dgrad[fused,H=384]", lambda: K.mlp_bwd_dgrad(dz, h16, w1, w2), r * (2 * d * 4 + 2 * h * 2), 4.0 * r * d * h),
            ("mlp_bwd_ln[fused,H=384,+mask]", lambda: K.mlp_bwd_ln(x, dout, w1, b1, w2, b2, gamma, want_mask=True),
             r * (3 * d * 4 + h * 2 + h // 8), 4.0 * r * d * h),
            ("mlp_bwd_ln[fused,H=384,mask only]", lambda: K.mlp_bwd_ln(x, dout, w1, b1, w2, b2, gamma, want_h=False, want_mask=True),
             r * (3 * d * 4 + h // 8), 4.0 * r * d * h),
            ("mlp_bwd_dgrad[fused,H=384,mask]", lambda: K.mlp_bwd_dgrad(dz, None, w1, w2, mask=mask), r * (2 * d * 4 + h * 2 + h // 8), 4.0 * r * d * h),
            ("mlp_bwd_dgrad[fused,H=384,mask,no dh]", lambda: K.mlp_bwd_dgrad(dz, None, w1, w2, mask=mask, want_dh=False),
             r * (2 * d * 4 + h // 8), 4.0 * r * d * h),
            ("attn_edge_fwd[fused,+a16]", lambda: K.attn_edge_fwd(x, q, k, w, b2, w, b2, gamma, beta, 0.25), r * d * 10, 4.0 * r * d * d),
            ("attn_edge_fwd[fused,+a16+e+z]", lambda: K.attn_edge_fwd(x, q, k, w, b2, w, b2, gamma, beta, 0.25, True, True, True),
             r * d * 18, 4.0 * r * d * d),
            ("softmax_agg16_fwd", lambda: K.softmax_agg16_fwd(a16, v), r * d * 2, 0.0),
            ("attn_edge forward pair [attn_edge_fwd + softmax_agg16]", fused_attn_edge, r * d * 8, 4.0 * r * d * d),
            ("attn_edge forward unfused [rows_gemm, attn_scores_fwd, rows_gemm, add_ln]", unfused_attn_edge, r * d * 8, 4.0 * r * d * d),
            ("attn_scores_fwd[fused]", lambda: K.attn_scores_fwd(q, k, v, x4, 0.25), 2 * r * d * 4, 0.0),
            ("attn_scores_fwd[stats only]", lambda: K.attn_scores_fwd(q, k, v, x4, 0.25, True, False), r * d * 4, 0.0),
            ("attn_scores_bwd[fused,stats]", lambda: K.attn_scores_bwd(dgn, dout.view(b, n, n, d), q, k, v, x4, 0.25, stats), 3 * r * d * 4, 0.0),
            ("attn_scores_bwd[fused,stats,de+=]", lambda: K.attn_scores_bwd(dgn, dout.view(b, n, n, d), q, k, v, x4, 0.25, stats, de_accum=acc4),
             4 * r * d * 4, 0.0),
            ("attn_scores_bwd[fused,stats,de16]", lambda: K.attn_scores_bwd(dgn, dout.view(b, n, n, d), q, k, v, x4, 0.25, stats, True),
             r * d * 10, 0.0),
            ("attn_scores_bwd[fused,stats,de16,da16]", lambda: K.attn_scores_bwd(dgn, a16.view(b, n, n, d), q, k, v, x4, 0.25, stats, True),
             r * d * 8, 0.0),
            ("rows_gemm[K=128,N=128,o16]", lambda: K.rows_gemm(x, w, False, out_bf16=True), r * d * 6, 2.0 * r * d * d),
            ("mlp_fwd[unfused: 2 rows_gemm + add_ln]", unfused, 2 * r * d * 4, 4.0 * r * d * h),
            ("rows_gemm[K=128,N=128]", lambda: K.rows_gemm(x, w, True, b2), 2 * r * d * 4, 2.0 * r * d * d),
            ("rows_gemm[K=128,N=128,+resid]", lambda: K.rows_gemm(x, w, True, b2, resid=dout), 3 * r * d * 4, 2.0 * r * d * d),
            ("rows_gemm[K=128,N=128,a16,+resid]", lambda: K.rows_gemm(a16, w, False, resid=dout), r * d * 10, 2.0 * r * d * d),
            ("rows_gemm[K=128,N=384,+relu]", lambda: K.rows_gemm(x, w1, True, b1, True), r * (d + h) * 4, 2.0 * r * d * h),
            ("gemm_tn[M=128,N=128]", lambda: K.gemm_tn(x, dout), 2 * r * d * 4, 2.0 * r * d * d),
            ("gemm_tn[M=128,N=128,b16]", lambda: K.gemm_tn(x, a16), r * d * 6, 2.0 * r * d * d),
            ("gemm_tn[M=384,N=128,a16]", lambda: K.gemm_tn(h16, x), r * (h * 2 + d * 4), 2.0 * r * d * h),
            ("gemm_tn[node rows R=92160,M=128,N=128]", lambda: K.gemm_tn(x[:92160], dout[:92160]), 2 * 92160 * d * 4, 2.0 * 92160 * d * d),
            ("rows_gemm[node rows R=92160,K=128,N=128]", lambda: K.rows_gemm(x[:92160], w, True, b2), 2 * 92160 * d * 4, 2.0 * 92160 * d * d),
            ("symmetrize", lambda: K.symmetrize(x4), 3 * r * d * 4, 0.0),
            ("add_ln_fwd", lambda: K.add_ln_fwd(x, dout, gamma, beta), 3 * r * d * 4, 0.0),
            ("add_ln_bwd", lambda: K.add_ln_bwd(x, dout, None, gamma), 3 * r * d * 4, 0.0),
            # second-order kernels of the gradient penalty (block_backward_backward)
            ("add_ln_bwd_bwd", lambda: K.add_ln_bwd_bwd(dout, None, None, x, x, None, gamma), 5 * r * d * 4, 0.0),
            ("modulate_bwd_bwd", lambda: K.modulate_bwd_bwd(q, k, x4, dout.view(b, n, n, d), q, k, x4, 0.25), 5 * r * d * 4, 0.0),
            ("softmax_agg_bwd_bwd", lambda: K.softmax_agg_bwd_bwd(dout.view(b, n, n, d), v, dgn, x4, v), 4 * r * d * 4, 0.0),
            ("softmax_agg_bwd", lambda: K.softmax_agg_bwd(dgn, x4, v), 3 * r * d * 4, 0.0),
            ("modulate_bwd", lambda: K.modulate_bwd(dout.view(b, n, n, d), q, k, x4, 0.25), 3 * r * d * 4, 0.0),
        ]
        for pf in ((_lib.PF_DEFAULT, args.pf) if args.pf is not None else (0, _lib.PF_ALL) if args.prefetch_ab else ((_lib.PF_DEFAULT, _lib.PF_DEFAULT | _lib.PF_CHAIN_KEEP) if args.keep_ab else (_lib.PF_DEFAULT,))):
            K.set_option(_lib.OPT_L2_PREFETCH, pf)
            for name, fn, nbytes, flops in benches:
                report(name, timeit(fn), nbytes, flops, l2_prefetch=pf)
        K.set_option(_lib.OPT_L2_PREFETCH, _lib.PF_DEFAULT)
        del dout, dz, h16, a16, x4
        del x
        # encoder-only forward (BASELINE config 5): molecules/s
        torch.manual_seed(0)
        enc = dg.TransformerEncoder(dim=d, depth=args.depth, heads=8, act=None, mlp_ratio=3, drop_rate=0.0).to(dev)
        xn, ye = rn(b, n, d), rn(b, n, n, d)
        with torch.no_grad():
            ms = timeit(lambda: enc(xn, ye), iters=5, warm=2)
        f_enc = d * d * (n * n * 16 + n * 20) * args.depth * b
        report(f"encoder_forward[L={args.depth},B={b},N={n}]", ms, args.depth * b * (2 * n * n * d + 2 * n * d) * 4, 2.0 * f_enc / 2,
               molecules_per_s=round(b / ms * 1e3, 1))


if __name__ == "__main__":
    main()
