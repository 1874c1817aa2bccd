#!/usr/bin/env python
"""Is the CPU arm that bench.py times (oracle/encoder_oracle.py, kind "port") as fast as the real thing?

Times, alternating, one GAN step (train.py:351-384) of (a) the UNMODIFIED reference modules + reference loss.py + AdamW
imported from /root/reference (build container only) and (b) the oracle port, same weights, same batch, same threads.
Prints both molecules/s; bench.py's CPU arm is honest when (b) >= (a) within noise."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("DRUGGEN_REFERENCE", "/root/reference")


def main():
    b, n, depth, rounds = int(os.environ.get("B", 8)), 45, int(os.environ.get("DEPTH", 8)), int(os.environ.get("ROUNDS", 3))
    threads = int(os.environ.get("THREADS", os.cpu_count() or 1))
    torch.set_num_threads(threads)
    from oracle import encoder_oracle as orc
    sys.path.insert(0, REF)
    from src.model.models import Generator, Discriminator          # the reference's own modules
    from src.model.loss import discriminator_loss, generator_loss  # and loss.py
    torch.manual_seed(0)
    G = Generator("relu", n, 5, 13, 0.0, dim=128, depth=depth, heads=8, mlp_ratio=3)
    D = Discriminator("relu", n, 5, 13, 0.0, dim=128, depth=depth, heads=8, mlp_ratio=3)
    g_opt, d_opt = torch.optim.AdamW(G.parameters(), 1e-5, (0.9, 0.999)), torch.optim.AdamW(D.parameters(), 1e-5, (0.9, 0.999))
    port = orc.OracleGAN(dict(G.state_dict()), dict(D.state_dict()), depth, depth, 8)
    a, x = orc.synthetic_batch(b, n, 13, 5, seed=1)

    def ref_step():                       # train.py:351-384
        g_opt.zero_grad(set_to_none=True); d_opt.zero_grad(set_to_none=True)
        _, _, d_loss = discriminator_loss(G, D, a, x, a, x, b, "cpu", 10.0)
        d_loss.item(); d_loss.backward(); d_opt.step()
        g_opt.zero_grad(set_to_none=True); d_opt.zero_grad(set_to_none=True)
        g_loss = generator_loss(G, D, a, x, b)[0]
        g_loss.item(); g_loss.backward(); g_opt.step()

    def port_step():
        port.step(a, x, a, x, torch.rand(b, 1, 1, 1), torch.rand(b, 1, 1))

    ref_step(); port_step()
    tr = tp = 0.0
    for _ in range(rounds):
        t0 = time.perf_counter(); ref_step(); tr += time.perf_counter() - t0
        t0 = time.perf_counter(); port_step(); tp += time.perf_counter() - t0
    print(json.dumps({"threads": threads, "batch": b, "depth": depth, "reference_mol_s": b * rounds / tr,
                      "port_mol_s": b * rounds / tp, "port_over_reference": tr / tp}))


if __name__ == "__main__":
    main()
