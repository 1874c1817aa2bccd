#!/usr/bin/env python
"""Debug aid: per block of a depth-8 Generator, the ATTN chain's outputs (y3, a16, E, z4) on the block's actual inputs vs emulation."""
import os, sys, math
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import druggen_b200 as dg
from druggen_b200 import kernels as K, block as blk
from emul_kernels import EmulBackend
from conftest import rel_l2
from oracle import encoder_oracle as orc

torch.manual_seed(21)
n, bsz, depth = 45, 2, 8
G = dg.Generator("relu", n, 5, 13, 0.0, dim=128, depth=depth, heads=8, mlp_ratio=3)
a, x = orc.synthetic_batch(bsz, n, 13, 5, seed=3)
dev = torch.device("cuda:0")
G.to(dev)
ins = []
for b_ in G.TransformerEncoder.Encoder_Blocks:
    b_.register_forward_pre_hook(lambda m, args: ins.append((args[0].detach().clone(), args[1].detach().clone())))
with dg.precision("bf16"), torch.no_grad():
    G(a.to(dev), x.to(dev))
    for li, (xi, yi) in enumerate(ins):
        p = {k: v.detach() for k, v in zip(blk.BLOCK_PARAM_NAMES, G.TransformerEncoder.Encoder_Blocks[li]._params())}
        d, c = 128, 1.0 / math.sqrt(128 // 8)
        def run(xi, yi, p):
            x1 = K.add_ln_fwd(xi.reshape(-1, d), None, p["ln1.weight"], p["ln1.bias"])
            q = K.rows_gemm(x1, p["attn.q.weight"], True, p["attn.q.bias"]).view(bsz, n, d)
            k = K.rows_gemm(x1, p["attn.k.weight"], True, p["attn.k.bias"]).view(bsz, n, d)
            outs = K.attn_edge_fwd(yi.reshape(-1, d), q, k, p["attn.e.weight"], p["attn.e.bias"], p["attn.out_e.weight"], p["attn.out_e.bias"],
                                   p["ln4.weight"], p["ln4.bias"], c, want_a16=True, want_e=True, want_z=True)
            return [o.float() for o in outs] + [q, k]
        got = run(xi, yi, p)
        torch.cuda.synchronize()
        K._install_backend_for_tests(EmulBackend(emulate_bf16=True))
        ref = run(xi.cpu(), yi.cpu(), {k_: v.cpu() for k_, v in p.items()})
        K._install_backend_for_tests(None)
        print(li, "y3 %.2e a16 %.2e E %.2e z4 %.2e q %.2e k %.2e" % tuple(rel_l2(g, r) for g, r in zip(got, ref)),
              " |E|max %.1f |a|max %.1f" % (float(ref[2].abs().max()), float(ref[1].abs().max())))
        q_, k_ = got[4], got[5]
        args = (yi.reshape(-1, d), q_, k_, p["attn.e.weight"], p["attn.e.bias"], p["attn.out_e.weight"], p["attn.out_e.bias"], p["ln4.weight"], p["ln4.bias"], c)
        o1 = K.attn_edge_fwd(*args, want_a16=True)
        o2 = K.attn_edge_fwd(*args, want_a16=False, want_e=True, want_z=True)
        o3 = K.attn_edge_fwd(*args, want_a16=False)
        print("    [+a16] y3 %.2e a16 %.2e | [+e+z] y3 %.2e E %.2e z %.2e | [] y3 %.2e" % (
            rel_l2(o1[0], got[0]), rel_l2(o1[1].float(), got[1]), rel_l2(o2[0], got[0]), rel_l2(o2[2], got[2]), rel_l2(o2[3], got[3]), rel_l2(o3[0], got[0])))
