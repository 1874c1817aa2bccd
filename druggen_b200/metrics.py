"""SNN / internal-diversity metric of the reference (src/util/utils.py:540-611) on bit-packed fingerprints.

``average_agg_tanimoto`` keeps the reference's signature.  The reference walks 5000 x 5000 blocks of an fp32 GEMM on 0/1 matrices
(``torch.mm`` counts the common bits) and aggregates on the host; here both sets are packed to 64-bit words once and one kernel
walks all pairs with AND + POPC (``dg_tanimoto_agg``): the pair counts are the same integers, the similarity is one IEEE fp32
division of them, so per-pair values -- and the 'max' aggregation -- are bit-equal to the reference's; the 'mean' aggregation sums
in fp32 per 128-fingerprint tile and fp64 across tiles (the reference: numpy fp32 pairwise sums per block, fp64 across blocks).
"""
from __future__ import annotations

import numpy as np
import torch

from . import kernels as K


def _packed(vecs, device):
    t = torch.as_tensor(np.ascontiguousarray(vecs))
    t = (t != 0).to(torch.uint8) if t.dtype != torch.uint8 else t
    return K.pack_bits(t.to(device).contiguous())


def average_agg_tanimoto(stock_vecs, gen_vecs, batch_size=5000, agg='max', device='cuda', p=1, intdiv=False):
    """src/util/utils.py:566-611.  ``batch_size`` is accepted for signature compatibility (the kernel does not block the
    pair matrix); ``device`` must be a CUDA device (no CPU fallback)."""
    assert agg in ['max', 'mean'], "Can aggregate only max or mean"
    stock, gen = _packed(stock_vecs, device), _packed(gen_vecs, device)
    n_stock = len(stock_vecs)
    if agg == 'max':
        out = K.tanimoto_agg(stock, gen, "max").cpu().numpy()                       # float32, exact per pair
        agg_tanimoto = (out ** p if p != 1 else out).astype(np.float64)             # (:598-599: x -> x^p is monotone: max commutes)
    else:
        agg_tanimoto = K.tanimoto_agg(stock, gen, "sum", p).cpu().numpy() / max(n_stock, 1)      # (:604-607)
    if p != 1:
        agg_tanimoto = (agg_tanimoto) ** (1 / p)
    if intdiv:
        return agg_tanimoto
    return np.mean(agg_tanimoto)


def internal_diversity(gen, device='cuda'):
    """src/util/utils.py:550-563 (``device``: where the fingerprints are packed and compared; the reference's is the CPU)."""
    diversity = [1 - x for x in average_agg_tanimoto(gen, gen, agg="mean", device=device, intdiv=True)]
    return np.mean(diversity), np.std(diversity)
