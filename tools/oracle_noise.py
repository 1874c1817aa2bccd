#!/usr/bin/env python
"""How well is the fp32 CPU oracle itself conditioned?  (build container, no GPU)

G-step gradients (depth 8, N = 45, B = 2) of the fp32 oracle against the fp64 oracle, single-threaded and multi-threaded.
Every individual ATen op is accurate to ~1e-7 in both settings, but a different summation order perturbs the forward by ~1e-7,
which flips the sign of a handful of the 25 M ReLU pre-activations (layers.py:52): the gradient is discontinuous there, so two
fp32 evaluations of the reference disagree by ~2e-3 rel-L2 on every parameter gradient.  Gradient error therefore scales like
sqrt(forward error): fp32 (1e-7) -> 5e-7..2e-3, bf16x3 (2e-6) -> 2e-3, bf16 (1e-3) -> 3e-2 -- which is what the GPU tests measure.
usage: python tools/oracle_noise.py <threads>"""
import sys, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import druggen_b200 as dg
from oracle import encoder_oracle as orc
from conftest import rel_l2
torch.set_num_threads(int(sys.argv[1]))
depth,n,bsz=8,45,2
torch.manual_seed(21)
G = dg.Generator("relu", n, 5, 13, 0.0, dim=128, depth=depth, heads=8, mlp_ratio=3)
D = dg.Discriminator("relu", n, 5, 13, 0.0, dim=128, depth=depth, heads=8, mlp_ratio=3)
a, x = orc.synthetic_batch(bsz, n, 13, 5, seed=3)
def run(dtype):
    ref = orc.OracleGAN({k:v.to(dtype) for k,v in G.state_dict().items()}, {k:v.to(dtype) for k,v in D.state_dict().items()}, depth, depth, 8)
    g_ref = ref.g_loss(a.to(dtype), x.to(dtype)); g_ref.backward()
    return g_ref.item(), {k: v.grad.clone() for k, v in ref.gp_.items()}
l64,g64=run(torch.float64); l32,g32=run(torch.float32)
print(l64,l32)
errs=sorted(((rel_l2(g32[k],g64[k]), k, float(g64[k].norm())) for k in g64), reverse=True)
for e in errs[:8]: print(e)
print('...'); 
for e in errs[-3:]: print(e)
import torch
allv=lambda g: torch.cat([g[k].double().flatten() for k in g64])
print('overall', rel_l2(allv(g32), allv(g64)))
