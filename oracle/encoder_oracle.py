"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Functional CPU restatement (plain PyTorch, fp32 or fp64, autograd for every
derivative) of the DrugGEN graph-transformer hot path.  Weights are addressed by the
reference's state-dict keys so a reference checkpoint drives it directly.

Reference lines restated (paths relative to the reference repo):
  block_forward          src/model/layers.py:174-193  (Encoder_Block.forward)
  attention part         src/model/layers.py:97-137   (MHA.forward)
  mlp part               src/model/layers.py:41-54    (MLP.forward)
  encoder_forward        src/model/layers.py:221-234  (TransformerEncoder.forward)
  generator_forward      src/model/models.py:71-103
  discriminator_forward  src/model/models.py:180-209
  gradient_penalty       src/model/loss.py:4-49
  discriminator_loss     src/model/loss.py:52-72
  generator_loss         src/model/loss.py:75-84
  gan_step               train.py:351-384
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

_ACTS = {
    "relu": torch.relu,
    "leaky": lambda t: torch.where(t > 0, t, 0.01 * t),
    "sigmoid": torch.sigmoid,
    "tanh": torch.tanh,
}


def _affine(t, p: Params, name: str):
    # nn.Linear.forward == F.linear (addmm), the op the reference modules dispatch to
    return F.linear(t, p[name + ".weight"], p[name + ".bias"])


def _layernorm(t, p: Params, name: str, eps: float = 1e-5):
    # nn.LayerNorm.forward == F.layer_norm (native_layer_norm), as in the reference
    return F.layer_norm(t, (t.shape[-1],), p[name + ".weight"], p[name + ".bias"], eps)


def block_forward(x, y, p: Params, prefix: str, heads: int):
    """One encoder block.  x:[B,N,d] nodes, y:[B,N,N,d] edges (layers.py:174-193)."""
    d = x.shape[-1]
    scale = 1.0 / math.sqrt(d // heads)            # layers.py:124 (d_k, not dim)
    x1 = _layernorm(x, p, prefix + "ln1")          # layers.py:185
    a = prefix + "attn."
    q = _affine(x1, p, a + "q")                    # layers.py:111-113
    k = _affine(x1, p, a + "k")
    v = _affine(x1, p, a + "v")
    e = _affine(y, p, a + "e")                     # layers.py:116
    # per-channel Hadamard "attention", no reduction over channels (layers.py:123-125)
    s = (q[:, :, None, :] * k[:, None, :, :]) * scale
    att = s * (e + 1.0) * e
    y1 = _affine(att, p, a + "out_e")              # layers.py:127, PRE-softmax scores
    prob = torch.softmax(att, dim=2)               # layers.py:130, over key atom j
    agg = (prob * v[:, None, :, :]).sum(dim=2)     # layers.py:132-134
    x2 = x1 + _affine(agg, p, a + "out_n")         # layers.py:135,187 (residual on x1)
    y2 = y + y1                                    # layers.py:188
    x3 = _layernorm(x2, p, prefix + "ln3")         # layers.py:189
    y3 = _layernorm(y2, p, prefix + "ln4")         # layers.py:190
    hx = torch.relu(_affine(x3, p, prefix + "mlp.fc1"))
    hy = torch.relu(_affine(y3, p, prefix + "mlp2.fc1"))
    x_out = _layernorm(x3 + _affine(hx, p, prefix + "mlp.fc2"), p, prefix + "ln5")   # :191
    y_out = _layernorm(y3 + _affine(hy, p, prefix + "mlp2.fc2"), p, prefix + "ln6")  # :192
    return x_out, y_out


def encoder_forward(x, y, p: Params, depth: int, heads: int, prefix: str = "Encoder_Blocks."):
    """TransformerEncoder.forward (layers.py:221-234)."""
    for layer in range(depth):
        x, y = block_forward(x, y, p, f"{prefix}{layer}.", heads)
    return x, y


def _prologue(z_e, z_n, p: Params, act: str):
    f = _ACTS[act]
    node = f(_affine(f(_affine(z_n, p, "node_layers.0")), p, "node_layers.2"))   # models.py:52-56,91
    edge = f(_affine(f(_affine(z_e, p, "edge_layers.0")), p, "edge_layers.2"))   # models.py:57-61,92
    edge = (edge + edge.permute(0, 2, 1, 3)) / 2                                 # models.py:94
    return node, edge


def generator_forward(z_e, z_n, p: Params, depth: int, heads: int, act: str = "relu"):
    """Generator.forward (models.py:71-103): returns (node, edge, node_sample, edge_sample)."""
    node, edge = _prologue(z_e, z_n, p, act)
    node, edge = encoder_forward(node, edge, p, depth, heads, "TransformerEncoder.Encoder_Blocks.")
    return node, edge, _affine(node, p, "readout_n"), _affine(edge, p, "readout_e")


def discriminator_forward(z_e, z_n, p: Params, depth: int, heads: int, act: str = "relu"):
    """Discriminator.forward (models.py:180-209): returns [B,1] logits."""
    f = _ACTS[act]
    node, edge = _prologue(z_e, z_n, p, act)
    node, _ = encoder_forward(node, edge, p, depth, heads, "TransformerEncoder.Encoder_Blocks.")
    h = node.reshape(node.shape[0], -1)
    h = f(_affine(h, p, "node_mlp.0"))
    h = f(_affine(h, p, "node_mlp.2"))
    h = f(_affine(h, p, "node_mlp.4"))
    return _affine(h, p, "node_mlp.6")


def gradient_penalty(d_fn, real_node, real_edge, fake_node, fake_edge, eps_edge, eps_node):
    """loss.py:4-49 with the two U[0,1) draws passed in (eps_edge drawn first, loss.py:21-22)."""
    int_node = (eps_node * real_node + (1 - eps_node) * fake_node).requires_grad_(True)
    int_edge = (eps_edge * real_edge + (1 - eps_edge) * fake_edge).requires_grad_(True)
    logits = d_fn(int_edge, int_node)
    g_node, g_edge = torch.autograd.grad(logits, [int_node, int_edge], torch.ones_like(logits),
                                         create_graph=True, retain_graph=True)
    b = real_node.shape[0]
    g = torch.cat([g_node.reshape(b, -1), g_edge.reshape(b, -1)], dim=1)
    return ((g.norm(2, dim=1) - 1) ** 2).mean()


def discriminator_loss(g_fn, d_fn, drug_adj, drug_annot, mol_adj, mol_annot, eps_edge, eps_node, lambda_gp):
    """loss.py:52-72."""
    real = -d_fn(drug_adj, drug_annot).mean()
    _, _, node_sample, edge_sample = g_fn(mol_adj, mol_annot)
    node_sample, edge_sample = node_sample.detach(), edge_sample.detach()
    fake = d_fn(edge_sample, node_sample).mean()
    gp = gradient_penalty(d_fn, drug_annot, drug_adj, node_sample, edge_sample, eps_edge, eps_node)
    return fake + real + lambda_gp * gp


def generator_loss(g_fn, d_fn, mol_adj, mol_annot):
    """loss.py:75-84."""
    _, _, node_sample, edge_sample = g_fn(mol_adj, mol_annot)
    return -d_fn(edge_sample, node_sample).mean()


class OracleGAN:
    """Weights + AdamW for the CPU GAN step (train.py:213-214,351-384).

    Used as (a) the parity checker in tests and (b) the CPU arm timed by
    ``bench.py`` (`cpu_baseline` and `--impl reference`, kind "port").
    """

    def __init__(self, g_state: Params, d_state: Params, depth: int, ddepth: int, heads: int,
                 act: str = "relu", lr: float = 1e-5, betas: Tuple[float, float] = (0.9, 0.999),
                 lambda_gp: float = 10.0):
        self.gp_ = {k: v.detach().clone().requires_grad_(True) for k, v in g_state.items()}
        self.dp_ = {k: v.detach().clone().requires_grad_(True) for k, v in d_state.items()}
        self.depth, self.ddepth, self.heads, self.act = depth, ddepth, heads, act
        self.lambda_gp = lambda_gp
        self.g_opt = torch.optim.AdamW(list(self.gp_.values()), lr, betas)   # train.py:213
        self.d_opt = torch.optim.AdamW(list(self.dp_.values()), lr, betas)   # train.py:214

    def G(self, z_e, z_n):
        return generator_forward(z_e, z_n, self.gp_, self.depth, self.heads, self.act)

    def D(self, z_e, z_n):
        return discriminator_forward(z_e, z_n, self.dp_, self.ddepth, self.heads, self.act)

    def _zero(self):
        for t in list(self.gp_.values()) + list(self.dp_.values()):
            t.grad = None

    def d_loss(self, drug_adj, drug_annot, mol_adj, mol_annot, eps_edge, eps_node):
        return discriminator_loss(self.G, self.D, drug_adj, drug_annot, mol_adj, mol_annot,
                                  eps_edge, eps_node, self.lambda_gp)

    def g_loss(self, mol_adj, mol_annot):
        return generator_loss(self.G, self.D, mol_adj, mol_annot)

    def step(self, drug_adj, drug_annot, mol_adj, mol_annot, eps_edge, eps_node, skip_dead_d_grads: bool = False):
        """One train.py:351-384 iteration.  Returns (d_loss, g_loss) as floats.
        ``skip_dead_d_grads``: do not compute the Discriminator weight gradients of the G-step, which train.py:371-377
        computes and reset_grad (train.py:352) discards unread -- the same switch as GANTrainer's, so that bench.py's two
        arms do the same work."""
        self._zero()
        d = self.d_loss(drug_adj, drug_annot, mol_adj, mol_annot, eps_edge, eps_node)
        d_val = d.item()
        d.backward()
        self.d_opt.step()
        self._zero()
        if skip_dead_d_grads:
            for t in self.dp_.values():
                t.requires_grad_(False)
        try:
            g = self.g_loss(mol_adj, mol_annot)
        finally:
            for t in self.dp_.values():
                t.requires_grad_(True)
        g_val = g.item()
        g.backward()
        self.g_opt.step()
        return d_val, g_val


def synthetic_batch(batch: int, n: int, m_dim: int = 13, b_dim: int = 5, seed: int = 1,
                    dtype=torch.float32):
    """Synthetic one-hot molecules in the layout load_molecules produces
    (src/data/utils.py:128-143): x[B,N,m] and symmetric zero-diagonal a[B,N,N,b]."""
    g = torch.Generator().manual_seed(seed)
    atoms = torch.randint(0, m_dim, (batch, n), generator=g)
    upper = torch.triu(torch.randint(0, b_dim, (batch, n, n), generator=g), diagonal=1)
    bonds = upper + upper.transpose(1, 2)
    x = torch.nn.functional.one_hot(atoms, m_dim).to(dtype)
    a = torch.nn.functional.one_hot(bonds, b_dim).to(dtype)
    return a, x
