#!/usr/bin/env python
"""Gradient error of a precision mode against the fp64 CPU oracle over several seeds of the depth-8, N = 45 GAN step
(tests/test_parity_gpu.py::test_gan_step_depth8_n45_vs_oracle, one seed there).  Small batches sit on ReLU kinks: a unit of the
Discriminator head whose pre-activation is ~1e-7 from zero flips with the last-bit rounding of ANY implementation (the fp32 CPU
reference included, profiles/r02_oracle_noise.json), which moves a 2-molecule gradient by tens of percent.

    python tools/parity_sweep.py [mode=bf16] [batch=2] [seeds=21,22,23,24]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import druggen_b200 as dg
from conftest import rel_l2
from oracle import encoder_oracle as orc

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
bsz = int(sys.argv[2]) if len(sys.argv) > 2 else 2
seeds = [int(s) for s in (sys.argv[3] if len(sys.argv) > 3 else "21,22,23,24").split(",")]
n, dev = 45, torch.device("cuda:0")
flat = lambda gs: torch.cat([g.double().flatten().cpu() for g in gs])  # noqa: E731
for seed in seeds:
    torch.manual_seed(seed)
    G = dg.Generator("relu", n, 5, 13, 0.0, dim=128, depth=8, heads=8, mlp_ratio=3)
    D = dg.Discriminator("relu", n, 5, 13, 0.0, dim=128, depth=8, heads=8, mlp_ratio=3)
    f64 = lambda t: t.double()  # noqa: E731
    ref = orc.OracleGAN({k: f64(v) for k, v in G.state_dict().items()}, {k: f64(v) for k, v in D.state_dict().items()}, 8, 8, 8)
    a, x = orc.synthetic_batch(bsz, n, 13, 5, seed=seed + 100)
    da, dx = orc.synthetic_batch(bsz, n, 13, 5, seed=seed + 200)
    eps_e, eps_n = torch.rand(bsz, 1, 1, 1), torch.rand(bsz, 1, 1)
    d_ref = ref.d_loss(f64(da), f64(dx), f64(a), f64(x), f64(eps_e), f64(eps_n))
    d_ref.backward()
    gD_ref = {k: v.grad.clone() for k, v in ref.dp_.items() if v.grad is not None and float(v.grad.abs().max()) > 1e-12}
    ref._zero()
    g_ref = ref.g_loss(f64(a), f64(x))
    g_ref.backward()
    gG_ref = {k: v.grad.clone() for k, v in ref.gp_.items()}
    G.to(dev), D.to(dev)
    to = lambda t: t.to(dev)  # noqa: E731
    with dg.precision(mode):
        d = orc.discriminator_loss(G, D, to(da), to(dx), to(a), to(x), to(eps_e), to(eps_n), 10.0)
        d.backward()
        gD = {k: v.grad.clone() for k, v in D.named_parameters() if v.grad is not None}
        G.zero_grad(set_to_none=True), D.zero_grad(set_to_none=True)
        g = orc.generator_loss(G, D, to(a), to(x))
        g.backward()
        gG = {k: v.grad.clone() for k, v in G.named_parameters()}
    print(json.dumps({"mode": mode, "seed": seed, "batch": bsz,
                      "d_loss_rel": abs(d.item() - d_ref.item()) / max(1.0, abs(d_ref.item())),
                      "g_loss_rel": abs(g.item() - g_ref.item()) / max(1.0, abs(g_ref.item())),
                      "D_grads_rel_l2_all": rel_l2(flat(gD[k] for k in gD_ref), flat(gD_ref.values())),
                      "G_grads_rel_l2_all": rel_l2(flat(gG[k] for k in gG_ref), flat(gG_ref.values()))}), flush=True)
