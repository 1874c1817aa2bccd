#!/usr/bin/env python
"""Debug aid: the depth-8 GAN step (bf16 mode) on the GPU kernels vs the bf16-emulating torch backend; per-tensor errors."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import druggen_b200 as dg
from druggen_b200 import kernels as K
from emul_kernels import EmulBackend
from conftest import rel_l2
from oracle import encoder_oracle as orc

depth = int(os.environ.get("DEPTH", 8))
torch.manual_seed(21)
n, bsz = 45, 2
G = dg.Generator("relu", n, 5, 13, 0.0, dim=128, depth=depth, heads=8, mlp_ratio=3)
D = dg.Discriminator("relu", n, 5, 13, 0.0, dim=128, depth=depth, heads=8, mlp_ratio=3)
a, x = orc.synthetic_batch(bsz, n, 13, 5, seed=3)
da, dx = orc.synthetic_batch(bsz, n, 13, 5, seed=4)
eps_e, eps_n = torch.rand(bsz, 1, 1, 1), torch.rand(bsz, 1, 1)


def run(dev):
    to = lambda t: t.to(dev)
    G.to(dev), D.to(dev)
    G.zero_grad(set_to_none=True), D.zero_grad(set_to_none=True)
    with dg.precision("bf16"):
        d = orc.discriminator_loss(G, D, to(da), to(dx), to(a), to(x), to(eps_e), to(eps_n), 10.0)
        d.backward()
        gD = {k: v.grad.detach().cpu().clone() for k, v in D.named_parameters() if v.grad is not None}
        G.zero_grad(set_to_none=True), D.zero_grad(set_to_none=True)
        g = orc.generator_loss(G, D, to(a), to(x))
        g.backward()
        gG = {k: v.grad.detach().cpu().clone() for k, v in G.named_parameters()}
    return d.item(), g.item(), gD, gG

reps = int(os.environ.get("REPS", 2))
gpu = [run(torch.device("cuda:0")) for _ in range(reps)]
K._install_backend_for_tests(EmulBackend(emulate_bf16=True))
ref = run(torch.device("cpu"))
K._install_backend_for_tests(None)
if K._debug_hold:
    bad = 0
    for rec in K._debug_hold:
        if rec[0] == "clone" and rec[1].is_cuda:
            if not torch.equal(rec[1], rec[2]):
                bad += 1
                print("  ATTN out changed after the launch: max diff", float((rec[1] - rec[2]).abs().max()))
            if not torch.equal(rec[3], rec[4]):
                bad += 1
                print("  ATTN input y changed after the launch: max diff", float((rec[3] - rec[4]).abs().max()))
    print("held ATTN launches:", sum(1 for r in K._debug_hold if r[0] == "clone"), "changed later:", bad)
for i, r in enumerate(gpu):
    print(f"run {i}: d {r[0]:.6f} (emul {ref[0]:.6f})  g {r[1]:.6f} (emul {ref[1]:.6f})")
    for name, got, want in (("D", r[2], ref[2]), ("G", r[3], ref[3])):
        errs = sorted(((rel_l2(got[k], want[k]), k) for k in want if float(want[k].abs().max()) > 1e-12), reverse=True)
        print("  ", name, "worst:", [(f"{e:.2e}", k) for e, k in errs[:6]])
if reps > 1:
    print("run-to-run (GPU 0 vs 1):", max(rel_l2(gpu[0][2][k], gpu[1][2][k]) for k in gpu[0][2]), max(rel_l2(gpu[0][3][k], gpu[1][3][k]) for k in gpu[0][3]))
