"""Fused multi-tensor AdamW on flat buckets (reference train.py:213-214 ``torch.optim.AdamW(params, lr, betas)``).

One network = one flat fp32 parameter buffer (every ``nn.Parameter`` becomes a view into it, so ``state_dict`` /
``load_state_dict`` / checkpoints are unchanged), one flat gradient bucket, two flat moment buffers.  Per step: one
``_foreach_copy_`` packs the gradients that exist into the bucket, the data-parallel all-reduce (one NCCL collective, the
same bucket -- parallel.py's reducer needs no pack / unpack of its own) averages it, and ONE kernel (``dg_adamw_flat``)
applies torch's AdamW formulas to every element.  Parameters whose gradient is ``None`` (the Discriminator's dead last-block
edge weights, models.py:202-207) are skipped exactly as torch skips them: no decay, no moment update, no step count.
"""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import kernels as K

# device-side segment record of dg_adamw_flat (include/druggen_b200.h)
SEG_DTYPE = np.dtype([("begin", "<i8"), ("end", "<i8"), ("bc1", "<f4"), ("bc2_sqrt", "<f4"), ("active", "<i4"), ("pad", "<i4")])


class FlatAdamW:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 1e-2, process_group: Optional[object] = None):
        self.params = [p for p in params]
        assert self.params, "no parameters"
        self.lr, self.betas, self.eps, self.weight_decay, self.pg = lr, betas, eps, weight_decay, process_group
        dev = self.params[0].device
        assert all(p.device == dev and p.dtype == torch.float32 for p in self.params), "one device, fp32 parameters"
        sizes = [p.numel() for p in self.params]
        self.offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        total = int(self.offsets[-1])
        self.flat_p = torch.empty(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o, n in zip(self.params, self.offsets[:-1], sizes):
                view = self.flat_p[int(o):int(o) + n].view_as(p)
                view.copy_(p)
                p.data = view                      # the module's parameter IS a slice of the flat buffer from here on
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self._g_views = [self.flat_g[int(o):int(o) + n].view_as(p) for p, o, n in zip(self.params, self.offsets[:-1], sizes)]
        self.steps = np.zeros(len(self.params), dtype=np.int64)
        self._segs_host = np.zeros(len(self.params), dtype=SEG_DTYPE)
        self._segs_host["begin"], self._segs_host["end"] = self.offsets[:-1], self.offsets[1:]
        nbytes = len(self.params) * SEG_DTYPE.itemsize
        self._segs_dev = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        # the table goes up with an asynchronous copy from a small ring of pinned staging buffers: no host wait per step
        self._stage = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(4)] if dev.type == "cuda" else None
        self._stage_ev = [None] * 4
        self._stage_i = 0
        self.write_back_grads = False    # True: after step() every live .grad holds the rank-averaged gradient (tests)

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            if set_to_none or p.grad is None:
                p.grad = None
            else:
                p.grad.zero_()

    def world(self) -> int:
        return dist.get_world_size(self.pg) if dist.is_available() and dist.is_initialized() else 1

    @torch.no_grad()
    def step(self) -> None:
        live = [i for i, p in enumerate(self.params) if p.grad is not None]
        if not live:
            return
        torch._foreach_copy_([self._g_views[i] for i in live], [self.params[i].grad for i in live])
        world = self.world()
        if world > 1:        # ONE collective per backward; the dead segments ride along untouched by the update
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.pg)
            self.flat_g.mul_(1.0 / world)
        self.steps[live] += 1
        b1, b2 = self.betas
        seg = self._segs_host
        seg["active"] = 0
        seg["active"][live] = 1
        t = np.maximum(self.steps, 1).astype(np.float64)
        seg["bc1"] = (1.0 - b1 ** t).astype(np.float32)
        seg["bc2_sqrt"] = np.sqrt(1.0 - b2 ** t).astype(np.float32)
        raw = torch.from_numpy(seg.view(np.uint8).reshape(-1))
        if self._stage is None:
            self._segs_dev.copy_(raw)
        else:
            i = self._stage_i = (self._stage_i + 1) % len(self._stage)
            if self._stage_ev[i] is not None:
                self._stage_ev[i].synchronize()
            self._stage[i].copy_(raw)
            self._segs_dev.copy_(self._stage[i], non_blocking=True)
            self._stage_ev[i] = torch.cuda.Event()
            self._stage_ev[i].record()
        K.adamw_flat(self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq, self._segs_dev, len(self.params), self.lr, b1, b2,
                     self.eps, self.weight_decay)
        if world > 1 and self.write_back_grads:
            torch._foreach_copy_([self.params[i].grad for i in live], [self._g_views[i] for i in live])
