#!/usr/bin/env python
"""BASELINE config 5: encoder-only forward micro-benchmark.  Batch sweep x N in {9,45,90} x heads in {4,8};
prints one JSON line per point: molecules/s, algorithmic HBM GB/s and TFLOP/s with their fractions of the
measured peaks (MEASURED_PEAKS.json).  Under torchrun each rank runs the same sweep on its own GPU and
rank 0 reports the sum (replicas; the forward has no cross-molecule dependency)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import druggen_b200 as dg  # noqa: E402
from druggen_b200 import parallel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="256,1024,4096,16384,32768")
    ap.add_argument("--atoms", default="9,45,90")
    ap.add_argument("--heads", default="4,8")
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--max-edge-gb", type=float, default=24.0)
    args = ap.parse_args()
    rank, world, local = parallel.init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pk = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    dg.set_precision(args.precision)
    d, r = 128, 3
    for n in [int(v) for v in args.atoms.split(",")]:
        for heads in [int(v) for v in args.heads.split(",")]:
            torch.manual_seed(0)
            enc = dg.TransformerEncoder(dim=d, depth=args.depth, heads=heads, act=None, mlp_ratio=r, drop_rate=0.0).to(dev)
            for b in [int(v) for v in args.batches.split(",")]:
                if b * n * n * d * 4 / 2 ** 30 > args.max_edge_gb:
                    continue
                x, y = torch.randn(b, n, d, device=dev), torch.randn(b, n, n, d, device=dev)
                with torch.no_grad():
                    for _ in range(2):
                        enc(x, y)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    iters = 3
                    e0.record()
                    for _ in range(iters):
                        enc(x, y)
                    e1.record()
                    torch.cuda.synchronize()
                ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
                if world > 1:
                    torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
                ms = ms.item()
                flops = d * d * (n * n * (4 + 4 * r) + n * (8 + 4 * r)) * args.depth * b * world
                nbytes = args.depth * b * world * (2 * n * n * d + 2 * n * d) * 4
                if rank == 0:
                    print(json.dumps({"config": "encoder-only forward", "n_gpus": world, "batch_per_gpu": b, "atoms": n, "heads": heads,
                                      "depth": args.depth, "precision": args.precision, "ms": round(ms, 3),
                                      "molecules_per_s": round(b * world / ms * 1e3, 1),
                                      "TFLOPs": round(flops / ms / 1e9, 1), "tensor_frac": round(flops / ms / 1e9 / (pk["bf16_tflops"] * world), 4),
                                      "GBps_algorithmic": round(nbytes / ms / 1e6, 1), "hbm_frac": round(nbytes / ms / 1e6 / (pk["hbm_gbs"] * world), 4)}),
                          flush=True)
                del x, y


if __name__ == "__main__":
    main()
