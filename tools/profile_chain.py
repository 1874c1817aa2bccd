#!/usr/bin/env python
"""Launch each tcgen05 chain kernel at the bench size (B=512, N=45) for an ncu capture:
    ncu --set full --clock-control none --import-source on -k regex:mlp_chain_tc -s 5 -c 5 -o gpurun_out/chain python tools/profile_chain.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import druggen_b200 as dg  # noqa: E402
from druggen_b200 import kernels as K  # noqa: E402

b, n = int(os.environ.get("B", 512)), 45
dev = torch.device("cuda:0")
r, d, h = b * n * n, 128, 384
g = torch.Generator().manual_seed(0)
rn = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)  # noqa: E731
x, dy = rn(r, d), rn(r, d)
w1, b1, w2, b2 = rn(h, d, sc=d ** -0.5), rn(h, sc=0.1), rn(d, h, sc=h ** -0.5), rn(d, sc=0.1)
wd = rn(d, d, sc=d ** -0.5)
gamma, beta = torch.ones(d, device=dev), torch.zeros(d, device=dev)
q, k = rn(b, n, d), rn(b, n, d)
with dg.precision("bf16"):
    for _ in range(2):
        K.mlp_fwd(x, w1, b1, w2, b2, gamma, beta)
        dz, hh, _, _ = K.mlp_bwd_ln(x, dy, w1, b1, w2, b2, gamma)
        K.mlp_bwd_dgrad(dz, hh, w1, w2)
        K.attn_edge_fwd(x, q, k, wd, b2, wd, b2, gamma, beta, 0.25)
        K.attn_edge_fwd(x, q, k, wd, b2, wd, b2, gamma, beta, 0.25, True, True, True)
torch.cuda.synchronize()
print("done")
