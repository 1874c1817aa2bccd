#!/usr/bin/env python
"""Phase-level cycle breakdown of the tcgen05 chain kernels (debug counters, dg_debug_chain_profile).

For each chain kernel: average cycles per 128-row tile spent by the epilogue (warps 0 and 4), the loader
(warp 8) and the MMA thread in each phase (mostly: waiting on which barrier vs doing what)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import druggen_b200 as dg  # noqa: E402
from druggen_b200 import _lib, kernels as K  # noqa: E402

EPI = ["chunk:pre", "wait hacc_full", "wait hb_empty", "tmem_ld", "chunk fence+arrive+spill", "chunk tail", "gather resid", "wait z_full",
       "final tmem+sum", "stats bar", "lnbwd pass1", "final scatter", "chunk math", "chunk reuse waits", "chunk st_block"]
LOAD = ["loop/prefetch", "wait x_empty", "load+convert+store"]
MMA = ["issue", "wait w_full", "wait x_full", "wait hacc_empty", "wait hb_full", "wait z_empty"]


def main():
    dev = torch.device("cuda:0")
    b, n, d, h = int(os.environ.get("B", 512)), 45, 128, 384
    r = b * n * n
    g = torch.Generator(device="cpu").manual_seed(0)
    rn = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)  # noqa: E731
    x, dout = rn(r, d), rn(r, d)
    w1, b1, w2, b2 = rn(h, d, sc=d ** -0.5), rn(h, sc=0.1), rn(d, h, sc=h ** -0.5), rn(d, sc=0.1)
    w = rn(d, d, sc=d ** -0.5)
    gamma, beta = torch.ones(d, device=dev), torch.zeros(d, device=dev)
    q, k = rn(b, n, d), rn(b, n, d)
    lib = _lib.load()
    buf = torch.zeros(148 * 64, dtype=torch.int64, device=dev)
    tiles_per_cta = (r + 127) // 128 / 148
    with dg.precision("bf16"):
        dz, h16, _, _ = K.mlp_bwd_ln(x, dout, w1, b1, w2, b2, gamma)
        runs = [
            ("mlp_fwd", lambda: K.mlp_fwd(x, w1, b1, w2, b2, gamma, beta)),
            ("mlp_bwd_ln", lambda: K.mlp_bwd_ln(x, dout, w1, b1, w2, b2, gamma)),
            ("mlp_bwd_dgrad", lambda: K.mlp_bwd_dgrad(dz, h16, w1, w2)),
            ("attn_edge_fwd+a16", lambda: K.attn_edge_fwd(x, q, k, w, b2, w, b2, gamma, beta, 0.25)),
            ("attn_edge_fwd+a16+e+z", lambda: K.attn_edge_fwd(x, q, k, w, b2, w, b2, gamma, beta, 0.25, True, True, True)),
        ]
        for name, fn in runs:
            fn()
            torch.cuda.synchronize()
            buf.zero_()
            lib.dg_debug_chain_profile(buf.data_ptr())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            lib.dg_debug_chain_profile(None)
            t = buf.view(148, 4, 16).double().mean(0) / tiles_per_cta
            rec = {"kernel": name, "ms": round(e0.elapsed_time(e1), 3), "cycles_per_tile_total": round(float(t[0].sum()), 0)}
            for role, names, row in (("epi_w0", EPI, t[0]), ("epi_w4", EPI, t[1]), ("loader", LOAD, t[2]), ("mma", MMA, t[3])):
                rec[role] = {nm: round(float(row[i]), 0) for i, nm in enumerate(names)}
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
