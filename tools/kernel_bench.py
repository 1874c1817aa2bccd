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

    with dg.precision("bf16"):
        ms = timeit(lambda: K.mlp_fwd(x, w1, b1, w2, b2, gamma, beta))
        report("mlp_fwd[fused,H=384]", ms, 2 * r * d * 4, 4.0 * r * d * h)

        def unfused():
            hh = K.rows_gemm(x, w1, True, b1, True)
            m = K.rows_gemm(hh, w2, True, b2)
            return K.add_ln_fwd(x, m, gamma, beta)
        ms = timeit(unfused)
        report("mlp_fwd[unfused: 2 rows_gemm + add_ln]", ms, 2 * r * d * 4, 4.0 * r * d * h)
        w = rn(d, d, sc=d ** -0.5)
        ms = timeit(lambda: K.rows_gemm(x, w, True, b2))
        report("rows_gemm[K=128,N=128]", ms, 2 * r * d * 4, 2.0 * r * d * d)
        ms = timeit(lambda: K.gemm_tn(x, x))
        report("gemm_tn[M=128,N=128]", ms, 2 * r * d * 4, 2.0 * r * d * d)
        ms = timeit(lambda: K.add_ln_fwd(x, x, gamma, beta))
        report("add_ln_fwd", ms, 3 * r * d * 4, 0.0)
        ms = timeit(lambda: K.add_ln_bwd(x, x, x, gamma))
        report("add_ln_bwd", ms, 4 * r * d * 4, 0.0)
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
