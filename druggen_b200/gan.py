"""The GAN step that calls the hot path (reference loss.py:4-84 and train.py:351-384), restated
as a small trainer so bench.py / tests can run "one training iteration" without the reference's
PyG / RDKit data pipeline.  The reference's own ``loss.py`` runs unchanged on our modules too
(INTEGRATION.md); this file only exists because the reference tree is absent on the GPU box.

Multi-GPU: one process per GPU, the molecule batch sharded by rank, one flat-bucket NCCL
all-reduce of the Discriminator grads after ``d_loss.backward()`` and one of the Generator grads
after ``g_loss.backward()`` (parallel.py).
"""
from __future__ import annotations

import contextlib
from typing import Optional

import torch

from . import kernels as K
from . import ops, parallel
from .block import keep_intermediates
from .optim import FlatAdamW


def gradient_penalty(D, real_node, real_edge, fake_node, fake_edge, batch_size, device, keep: bool = False):
    """WGAN-GP term, loss.py:4-49 (eps_edge is drawn before eps_node, as there).  The glue around the double backward runs in
    three small kernels: the interpolation straight from the real side's labels when those are what the
    caller holds (``dg_gp_interp``, bit-identical to loss.py:21-26 on the one-hot tensor), and the per-sample norm /
    ``mean((|g| - 1)^2)`` with its gradient (``dg_gp_penalty``, ``dg_gp_penalty_bwd``: loss.py:42-47).
    ``keep``: the blocks of this pass may keep their forward intermediates (block.keep_intermediates) for the input-gradient
    pass and the second-order pass."""
    eps_edge = torch.rand(batch_size, 1, 1, 1, device=device)
    eps_node = torch.rand(batch_size, 1, 1, device=device)
    if torch.is_floating_point(real_node):
        int_node = (eps_node * real_node + (1 - eps_node) * fake_node).requires_grad_(True)
        int_edge = (eps_edge * real_edge + (1 - eps_edge) * fake_edge).requires_grad_(True)
    else:
        int_node = K.gp_interp(real_node, fake_node.contiguous(), eps_node).requires_grad_(True)
        int_edge = K.gp_interp(real_edge, fake_edge.contiguous(), eps_edge).requires_grad_(True)
    with keep_intermediates(keep):
        logits = D(int_edge, int_node)
    g_node, g_edge = torch.autograd.grad(logits, [int_node, int_edge], torch.ones_like(logits),
                                         create_graph=True, retain_graph=True)
    return ops.GradPenalty.apply(g_node, g_edge)


def discriminator_loss(G, D, drug_adj, drug_annot, mol_adj, mol_annot, batch_size, device, lambda_gp):
    """loss.py:52-72 -> (node, edge, d_loss)."""
    real = -D(drug_adj, drug_annot).mean()
    node, edge, node_sample, edge_sample = G(mol_adj, mol_annot)
    node_sample, edge_sample = node_sample.detach(), edge_sample.detach()
    fake = D(edge_sample, node_sample).mean()
    gp = gradient_penalty(D, drug_annot, drug_adj, node_sample, edge_sample, batch_size, device)
    return node, edge, fake + real + lambda_gp * gp


def generator_loss(G, D, mol_adj, mol_annot, batch_size):
    """loss.py:75-84 -> (g_loss, node, edge, node_sample, edge_sample)."""
    node, edge, node_sample, edge_sample = G(mol_adj, mol_annot)
    return -D(edge_sample, node_sample).mean(), node, edge, node_sample, edge_sample


def synthetic_molecules(batch: int, n: int, m_dim: int = 13, b_dim: int = 5, seed: int = 1, device="cpu", labels: bool = False):
    """Synthetic one-hot molecules in the layout ``load_molecules`` produces (src/data/utils.py:128-143):
    a[B,N,N,b] symmetric with a zero (class 0) diagonal, x[B,N,m]; fp32.  ``labels=True``: the same molecules in the
    label wire format -- uint8 bond labels [B,N,N] and atom labels [B,N] (what ``to_dense_adj`` holds before
    ``label2onehot``, src/data/utils.py:130-141), 1 byte per edge instead of 4 * b_dim."""
    g = torch.Generator().manual_seed(seed)
    atoms = torch.randint(0, m_dim, (batch, n), generator=g)
    upper = torch.triu(torch.randint(0, b_dim, (batch, n, n), generator=g), diagonal=1)
    bonds = upper + upper.transpose(1, 2)
    if labels:
        return bonds.to(torch.uint8).to(device), atoms.to(torch.uint8).to(device)
    x = torch.nn.functional.one_hot(atoms, m_dim).float()
    a = torch.nn.functional.one_hot(bonds, b_dim).float()
    return a.to(device), x.to(device)


@contextlib.contextmanager
def frozen(module):
    """``requires_grad_(False)`` on a module's parameters for the duration of the block."""
    ps = [p for p in module.parameters() if p.requires_grad]
    for p in ps:
        p.requires_grad_(False)
    try:
        yield
    finally:
        for p in ps:
            p.requires_grad_(True)


class GANTrainer:
    """Generator + Discriminator + two AdamW optimizers, stepped as train.py:351-384."""

    def __init__(self, G, D, lr_g: float = 1e-5, lr_d: float = 1e-5, betas=(0.9, 0.999), lambda_gp: float = 10.0,
                 process_group: Optional[object] = None, skip_dead_d_grads: bool = True, sequenced: bool = True):
        self.G, self.D, self.lambda_gp = G, D, lambda_gp
        # train.py:353-359 builds d_loss = fake + real + lambda * gp and calls ONE backward, so the graphs of all three
        # Discriminator passes are alive at once.  ``sequenced`` backpropagates the three terms one after the other into the same
        # .grad buffers (the gradient of a sum is the sum of the gradients; the accumulation order differs in the last fp32 bit):
        # a pass's activations die before the next pass starts, and the memory that frees lets the blocks of the plain passes keep
        # their forward intermediates instead of recomputing them in the backward (block.keep_intermediates).  The Generator's
        # forward inside the D step runs without a graph (loss.py:57-58 detaches its outputs anyway).
        self.sequenced = sequenced
        # train.py:371-377: g_loss.backward() also fills D's .grad, which nothing consumes (reset_grad, train.py:352,
        # zeroes it before the next D step; d_optimizer is not stepped).  With skip_dead_d_grads the Discriminator is
        # frozen while the G-step graph is built: the gradient still flows THROUGH D to G (dgrad), D's weight-gradient
        # contractions are not launched.  G's update is bit-identical either way (SURVEY 8d "necessary FLOPs").
        self.skip_dead_d_grads = skip_dead_d_grads
        # train.py:213-214 AdamW, as ONE fused launch per network over flat buckets; the data-parallel all-reduce (one per
        # backward) runs on the optimizer's own gradient bucket (optim.FlatAdamW)
        self.g_optimizer = FlatAdamW(G.parameters(), lr_g, betas, process_group=process_group)
        self.d_optimizer = FlatAdamW(D.parameters(), lr_d, betas, process_group=process_group)
        self.pg = process_group

    def reset_grad(self):
        self.g_optimizer.zero_grad(set_to_none=True)
        self.d_optimizer.zero_grad(set_to_none=True)

    # ---- checkpoint I/O, train.py:250-263: the reference's file names and contents (plain state_dicts with the reference's keys),
    # so checkpoints move between the two implementations in both directions
    def save_model(self, model_directory, idx, i):
        import os
        torch.save(self.G.state_dict(), os.path.join(model_directory, '{}-{}-G.ckpt'.format(idx + 1, i + 1)))
        torch.save(self.D.state_dict(), os.path.join(model_directory, '{}-{}-D.ckpt'.format(idx + 1, i + 1)))

    def restore_model(self, epoch, iteration, model_directory):
        import os
        for net, tag in ((self.G, "G"), (self.D, "D")):
            path = os.path.join(model_directory, '{}-{}-{}.ckpt'.format(epoch, iteration, tag))
            net.load_state_dict(torch.load(path, map_location=lambda storage, loc: storage))
        # the flat AdamW buckets alias the parameters: load_state_dict copies in place, nothing to rebuild

    def step(self, drug_adj, drug_annot, mol_adj, mol_annot):
        """One iteration on this rank's shard; returns (d_loss, g_loss) as Python floats
        (the two ``.item()`` syncs of train.py:364,380 included).  The four tensors are the reference's fp32 one-hots
        (``load_molecules``' output) or, equivalently, integer labels [B,N,N] / [B,N] -- the 1-byte wire format."""
        bsz, dev = mol_annot.shape[0], mol_annot.device
        self.reset_grad()
        if self.sequenced:
            d_val = self._d_step_sequenced(drug_adj, drug_annot, mol_adj, mol_annot, bsz, dev)
        else:
            _, _, d_loss = discriminator_loss(self.G, self.D, drug_adj, drug_annot, mol_adj, mol_annot, bsz, dev, self.lambda_gp)
            d_val = d_loss.item()
            d_loss.backward()
        self.d_optimizer.step()                 # (all-reduce of the D gradients inside, world > 1)
        self.reset_grad()
        with keep_intermediates(self.sequenced):
            with (frozen(self.D) if self.skip_dead_d_grads else contextlib.nullcontext()):
                g_loss = generator_loss(self.G, self.D, mol_adj, mol_annot, bsz)[0]
            g_val = g_loss.item()
            g_loss.backward()
        self.g_optimizer.step()                 # (all-reduce of the G gradients inside, world > 1)
        return d_val, g_val

    def _d_step_sequenced(self, drug_adj, drug_annot, mol_adj, mol_annot, bsz, dev) -> float:
        """loss.py:52-72 + ``d_loss.backward()`` term by term, in the reference's order of evaluation (real, G, fake, GP: the
        GP's eps draws are the only random numbers of the step).  Returns d_loss as a Python float (train.py:364)."""
        with keep_intermediates():
            real = -self.D(drug_adj, drug_annot).mean()
            real.backward()
            with torch.no_grad():
                _, _, node_sample, edge_sample = self.G(mol_adj, mol_annot)
            fake = self.D(edge_sample, node_sample).mean()
            fake.backward()
        gp = gradient_penalty(self.D, drug_annot, drug_adj, node_sample, edge_sample, bsz, dev, keep=True)
        d_val = (fake.detach() + real.detach() + self.lambda_gp * gp.detach()).item()     # (the sync of train.py:364, before the
        (self.lambda_gp * gp).backward()                                                  #  second-order pass is queued)
        return d_val
