#!/usr/bin/env python
"""Debug aid: one encoder block's first-order backward (and forward) on the GPU kernels vs the torch emulation of the same
launch sequence (tests/emul_kernels.py) at a given shape; prints rel-L2 per output."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import druggen_b200 as dg
from druggen_b200 import kernels as K, block as blk
from emul_kernels import EmulBackend
from conftest import rel_l2

b, n, d, heads = int(os.environ.get("B", 2)), int(os.environ.get("N", 45)), 128, 8
torch.manual_seed(0)
enc = dg.TransformerEncoder(dim=d, depth=1, heads=heads, act=None, mlp_ratio=3, drop_rate=0.0)
params = [p.detach() for p in enc.Encoder_Blocks[0]._params()]
x, y = torch.randn(b, n, d), torch.randn(b, n, n, d)
dxo, dyo = torch.randn(b, n, d), torch.randn(b, n, n, d)
dev = torch.device("cuda:0")
for want in (True, False):
    with dg.precision("bf16"):
        outs = blk.block_backward(x.to(dev), y.to(dev), dxo.to(dev), dyo.to(dev), [p.to(dev) for p in params], heads, True, want)
        fwd = blk.block_forward_nograd(x.to(dev), y.to(dev), [p.to(dev) for p in params], heads, True)
        torch.cuda.synchronize()
        K._install_backend_for_tests(EmulBackend(emulate_bf16=True))
        ref = blk.block_backward(x, y, dxo, dyo, params, heads, True, want)
        rfwd = blk.block_forward_nograd(x, y, params, heads, True)
        K._install_backend_for_tests(None)
    print("want_params", want, "fwd", rel_l2(fwd[0], rfwd[0]), rel_l2(fwd[1], rfwd[1]), "dx", rel_l2(outs[0], ref[0]), "dy", rel_l2(outs[1], ref[1]))
    for nm, g, r in zip(blk.BLOCK_PARAM_NAMES, outs[2], ref[2]):
        if g is not None:
            e = rel_l2(g, r)
            print(f"   {nm:22s} {e:.3e}" + ("   <<<<<<" if e > 2e-2 else ""))
ux, uy = torch.randn(b, n, d), torch.randn(b, n, n, d)
for edge_out in (True, False):
    with dg.precision("bf16"):
        o = blk.block_backward_backward(x.to(dev), y.to(dev), dxo.to(dev), dyo.to(dev) if edge_out else None, ux.to(dev), uy.to(dev),
                                        [p.to(dev) for p in params], heads, edge_out)
        torch.cuda.synchronize()
        K._install_backend_for_tests(EmulBackend(emulate_bf16=True))
        r = blk.block_backward_backward(x, y, dxo, dyo if edge_out else None, ux, uy, params, heads, edge_out)
        K._install_backend_for_tests(None)
    print("second order, edge_out", edge_out, "c_x", rel_l2(o[0], r[0]), "c_y", rel_l2(o[1], r[1]), "c_dxo", rel_l2(o[2], r[2]),
          "c_dyo", rel_l2(o[3], r[3]) if edge_out else None)
    for nm, g, rr in zip(blk.BLOCK_PARAM_NAMES, o[4], r[4]):
        if g is not None:
            e = rel_l2(g, rr)
            print(f"   {nm:22s} {e:.3e}" + ("   <<<<<<" if e > 2e-2 else ""))
