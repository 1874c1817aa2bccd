"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the UNMODIFIED
reference modules (imported from /root/reference, never copied) on seeded inputs.

Run in the build container (the GPU box has no /root/reference):
    python oracle/make_golden.py

Fixtures written:
  enc_fwd_cfg1.npz   BASELINE config 1: 1-layer TransformerEncoder forward, N=9 (B=8 stored)
  enc_grad.npz       2-layer encoder, forward + all first-order grads
  gan_step.npz       Generator+Discriminator (depth 1), loss.py D/G losses incl. the
                     gradient-penalty double backward, all grads, argmax decode
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("DRUGGEN_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from src.model.layers import TransformerEncoder  # noqa: E402  (reference, unmodified)
from src.model.models import Generator, Discriminator  # noqa: E402
from src.model import loss as ref_loss  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.encoder_oracle import synthetic_batch  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
torch.set_num_threads(4)


def sd_np(module, prefix="w::"):
    return {prefix + k: v.detach().numpy().copy() for k, v in module.state_dict().items()}


def grads_np(module, prefix="g::"):
    return {prefix + k: (p.grad.detach().numpy().copy() if p.grad is not None
                         else np.zeros(tuple(p.shape), np.float32))
            for k, p in module.named_parameters()}


def randomise_ln(module, gen):
    """LayerNorm affine params default to (1, 0); perturb so their grads/paths are exercised."""
    for m in module.modules():
        if isinstance(m, torch.nn.LayerNorm):
            with torch.no_grad():
                m.weight.add_(0.1 * torch.randn(m.weight.shape, generator=gen))
                m.bias.add_(0.1 * torch.randn(m.bias.shape, generator=gen))


def enc_fwd_cfg1():
    torch.manual_seed(0)
    enc = TransformerEncoder(dim=128, depth=1, heads=8, act=torch.nn.ReLU(), mlp_ratio=3, drop_rate=0.0)
    gen = torch.Generator().manual_seed(11)
    randomise_ln(enc, gen)
    x = torch.randn(8, 9, 128, generator=gen)
    y = torch.randn(8, 9, 9, 128, generator=gen)
    with torch.no_grad():
        xo, yo = enc(x, y)
    np.savez_compressed(os.path.join(OUT, "enc_fwd_cfg1.npz"), x=x.numpy(), y=y.numpy(),
                        x_out=xo.numpy(), y_out=yo.numpy(), depth=1, heads=8, mlp_ratio=3, **sd_np(enc))


def enc_grad():
    torch.manual_seed(1)
    enc = TransformerEncoder(dim=128, depth=2, heads=4, act=torch.nn.ReLU(), mlp_ratio=3, drop_rate=0.0)
    gen = torch.Generator().manual_seed(12)
    randomise_ln(enc, gen)
    x = torch.randn(3, 9, 128, generator=gen).requires_grad_(True)
    y = torch.randn(3, 9, 9, 128, generator=gen).requires_grad_(True)
    wx = torch.randn(3, 9, 128, generator=gen)
    wy = torch.randn(3, 9, 9, 128, generator=gen)
    xo, yo = enc(x, y)
    ((xo * wx).sum() + (yo * wy).sum()).backward()
    np.savez_compressed(os.path.join(OUT, "enc_grad.npz"), x=x.detach().numpy(), y=y.detach().numpy(),
                        wx=wx.numpy(), wy=wy.numpy(), x_out=xo.detach().numpy(), y_out=yo.detach().numpy(),
                        dx=x.grad.numpy(), dy=y.grad.numpy(), depth=2, heads=4, mlp_ratio=3,
                        **sd_np(enc), **grads_np(enc))


def gan_step():
    n, m_dim, b_dim, bsz, depth = 9, 13, 5, 4, 1
    torch.manual_seed(2)
    G = Generator("relu", n, b_dim, m_dim, 0.0, dim=128, depth=depth, heads=8, mlp_ratio=3)
    D = Discriminator("relu", n, b_dim, m_dim, 0.0, dim=128, depth=depth, heads=8, mlp_ratio=3)
    gen = torch.Generator().manual_seed(13)
    randomise_ln(G, gen)
    randomise_ln(D, gen)
    mol_a, mol_x = synthetic_batch(bsz, n, m_dim, b_dim, seed=21)
    drug_a, drug_x = synthetic_batch(bsz, n, m_dim, b_dim, seed=22)
    out = dict(mol_a=mol_a.numpy(), mol_x=mol_x.numpy(), drug_a=drug_a.numpy(), drug_x=drug_x.numpy(),
               depth=depth, heads=8, mlp_ratio=3, n=n, m_dim=m_dim, b_dim=b_dim, lambda_gp=10.0)
    out.update(sd_np(G, "wG::"))
    out.update(sd_np(D, "wD::"))

    # the two eps draws inside reference gradient_penalty (loss.py:21-22): edge first, then node
    torch.manual_seed(1234)
    eps_edge = torch.rand(bsz, 1, 1, 1)
    eps_node = torch.rand(bsz, 1, 1)
    out.update(eps_edge=eps_edge.numpy(), eps_node=eps_node.numpy())

    with torch.no_grad():
        node, edge, node_sample, edge_sample = G(mol_a, mol_x)
        out.update(G_node=node.numpy(), G_edge=edge.numpy(), G_node_sample=node_sample.numpy(),
                   G_edge_sample=edge_sample.numpy(), D_real=D(drug_a, drug_x).numpy(),
                   D_fake=D(edge_sample, node_sample).numpy(),
                   node_argmax=node_sample.argmax(-1).numpy(), edge_argmax=edge_sample.argmax(-1).numpy())
        for name, t in (("node", node_sample), ("edge", edge_sample)):
            top2 = t.topk(2, dim=-1).values
            out[f"{name}_gap"] = (top2[..., 0] - top2[..., 1]).numpy()

    # gradient penalty alone (value + D grads through the double backward)
    torch.manual_seed(1234)
    gp = ref_loss.gradient_penalty(D, drug_x, drug_a, node_sample, edge_sample, bsz, "cpu")
    D.zero_grad()
    gp.backward()
    out.update(gp=gp.item())
    out.update(grads_np(D, "gGP_D::"))

    # discriminator step (train.py:352-366, no optimizer step: grads are the golden)
    G.zero_grad(); D.zero_grad()
    torch.manual_seed(1234)
    _, _, d_loss = ref_loss.discriminator_loss(G, D, drug_a, drug_x, mol_a, mol_x, bsz, "cpu", 10.0)
    d_loss.backward()
    out.update(d_loss=d_loss.item())
    out.update(grads_np(D, "gD_D::"))

    # generator step (train.py:370-382) on the same (un-stepped) weights
    G.zero_grad(); D.zero_grad()
    g_loss = ref_loss.generator_loss(G, D, mol_a, mol_x, bsz)[0]
    g_loss.backward()
    out.update(g_loss=g_loss.item())
    out.update(grads_np(G, "gG_G::"))
    out.update(grads_np(D, "gG_D::"))
    np.savez_compressed(os.path.join(OUT, "gan_step.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    enc_fwd_cfg1()
    enc_grad()
    gan_step()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
