#!/usr/bin/env python
"""Launch one hot-path kernel a few times at the bench size so ncu can capture it:
    ncu --set full --clock-control none --import-source on -k regex:<name> -s 2 -c 1 -o gpurun_out/prof python tools/profile_one.py mlp_fwd
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import druggen_b200 as dg  # noqa: E402
from druggen_b200 import kernels as K  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "mlp_fwd"
b = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda:0")
r, d, h = b * 45 * 45, 128, 384
g = torch.Generator().manual_seed(0)
rn = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)  # noqa: E731
x = rn(r, d)
w1, b1, w2, b2 = rn(h, d, sc=d ** -0.5), rn(h, sc=0.1), rn(d, h, sc=h ** -0.5), rn(d, sc=0.1)
gamma, beta = torch.ones(d, device=dev), torch.zeros(d, device=dev)
with dg.precision("bf16"):
    for _ in range(4):
        if which == "mlp_fwd":
            K.mlp_fwd(x, w1, b1, w2, b2, gamma, beta)
        elif which == "rows_gemm":
            K.rows_gemm(x, w1, True, b1, True)
        elif which == "gemm_tn":
            K.gemm_tn(x, x)
torch.cuda.synchronize()
print("done")
