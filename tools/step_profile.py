#!/usr/bin/env python
"""torch.profiler view of one GAN step: which CUDA kernels that are NOT ours (ATen elementwise, cuBLAS, optimizer) take time."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import druggen_b200 as dg  # noqa: E402
from druggen_b200 import gan  # noqa: E402

bsz, n = int(os.environ.get("B", 1024)), 45
dev = torch.device("cuda:0")
dg.set_precision("bf16")
torch.manual_seed(0)
G = dg.Generator("relu", n, 5, 13, 0.0, dim=128, depth=8, heads=8, mlp_ratio=3).to(dev)
D = dg.Discriminator("relu", n, 5, 13, 0.0, dim=128, depth=8, heads=8, mlp_ratio=3).to(dev)
tr = gan.GANTrainer(G, D)
a, x = gan.synthetic_molecules(bsz, n, 13, 5, seed=1, device=dev, labels=True)
da, dx = gan.synthetic_molecules(bsz, n, 13, 5, seed=2, device=dev, labels=True)
for _ in range(2):
    tr.step(da, dx, a, x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.step(da, dx, a, x)
    torch.cuda.synchronize()
ev = prof.key_averages()
rows = sorted(((e.key, e.device_time_total / 1e3, e.count) for e in ev if e.device_time_total > 0), key=lambda r: -r[1])
tot = sum(r[1] for r in rows if not r[0].startswith("aten::") and not r[0].startswith("autograd") and "Backward" not in r[0])
print("kernels (ms, count):")
for k, ms, c in rows[:70]:
    if k.startswith("aten::") or "Backward" in k or k.startswith("autograd") or k.startswith("Optimizer") or k[0].isupper() and "Fn" in k:
        continue
    print(f"{ms:9.2f} {c:6d}  {k[:130]}")
