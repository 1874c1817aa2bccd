"""Data-parallel path on CPU: world_size=2 over gloo.  Each rank runs the generator step
(train.py:370-382) on its shard of the molecule batch; after ONE flat all-reduce the averaged
gradients equal the single-process gradients of the whole batch, and parameters with no
gradient (the Discriminator's dead last-block edge weights) stay None on every rank."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    import druggen_b200 as dg
    torch.manual_seed(0)
    G = dg.Generator("relu", 5, 5, 13, 0.0, dim=32, depth=2, heads=4, mlp_ratio=3)
    D = dg.Discriminator("relu", 5, 5, 13, 0.0, dim=32, depth=2, heads=4, mlp_ratio=3)
    return G, D


def _grads(G, D, a, x):
    from druggen_b200 import gan
    G.zero_grad(set_to_none=True); D.zero_grad(set_to_none=True)
    gan.generator_loss(G, D, a, x, a.shape[0])[0].backward()


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from druggen_b200 import gan, kernels, parallel
    from emul_kernels import EmulBackend
    kernels._install_backend_for_tests(EmulBackend())
    kernels.set_precision("fp32")
    torch.set_num_threads(1)
    parallel.init_from_env("gloo")
    G, D = _build()
    a, x = gan.synthetic_molecules(8, 5, 13, 5, seed=7)
    _grads(G, D, parallel.shard_batch(a, rank, world), parallel.shard_batch(x, rank, world))
    parallel.FlatGradReducer(G.parameters()).all_reduce_mean()
    parallel.FlatGradReducer(D.parameters()).all_reduce_mean()
    torch.save({"G": [p.grad for p in G.parameters()], "D": [p.grad for p in D.parameters()]},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_matches_single_process(tmp_path):
    port = 29611 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from druggen_b200 import gan, kernels
    from emul_kernels import EmulBackend
    kernels._install_backend_for_tests(EmulBackend())
    kernels.set_precision("fp32")
    try:
        G, D = _build()
        a, x = gan.synthetic_molecules(8, 5, 13, 5, seed=7)
        _grads(G, D, a, x)
    finally:
        kernels._install_backend_for_tests(None)
    r0, r1 = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    for net, mod in (("G", G), ("D", D)):
        for p, g0, g1 in zip(mod.parameters(), r0[net], r1[net]):
            if p.grad is None:
                assert g0 is None and g1 is None
                continue
            assert torch.equal(g0, g1)                       # every rank holds the same averaged gradient
            assert torch.allclose(g0, p.grad, rtol=2e-4, atol=1e-6)
    assert any(g is None for g in r0["D"])                  # dead D edge parameters stay None


def _trainer_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from druggen_b200 import gan, kernels, parallel
    from emul_kernels import EmulBackend
    kernels._install_backend_for_tests(EmulBackend())
    kernels.set_precision("fp32")
    torch.set_num_threads(1)
    parallel.init_from_env("gloo")
    G, D = _build()
    tr = gan.GANTrainer(G, D, lr_g=1e-3, lr_d=1e-3, lambda_gp=0.0, process_group=dist.group.WORLD)
    tr.g_optimizer.write_back_grads = True
    a, x = gan.synthetic_molecules(8, 5, 13, 5, seed=7)
    da, dx = gan.synthetic_molecules(8, 5, 13, 5, seed=8)
    sh = lambda t: parallel.shard_batch(t, rank, world)  # noqa: E731
    losses = tr.step(sh(da), sh(dx), sh(a), sh(x))
    torch.save({"losses": losses, "G": [p.detach().clone() for p in G.parameters()], "D": [p.detach().clone() for p in D.parameters()],
                "gG": [None if p.grad is None else p.grad.clone() for p in G.parameters()]}, os.path.join(out_dir, f"tr{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_trainer_step_matches_single_process(tmp_path):
    """GANTrainer(process_group=...) over gloo, world 2: the sequenced D step accumulates its three backward passes locally and the
    optimizer all-reduces ONE flat bucket per network (optim.FlatAdamW); after a full step both ranks hold the same weights, equal
    to a single process stepping on the whole batch (lambda_gp = 0: the penalty's eps draws are per rank by design)."""
    port = 29811 + os.getpid() % 150
    mp.spawn(_trainer_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from druggen_b200 import gan, kernels
    from emul_kernels import EmulBackend
    kernels._install_backend_for_tests(EmulBackend())
    kernels.set_precision("fp32")
    try:
        G, D = _build()
        tr = gan.GANTrainer(G, D, lr_g=1e-3, lr_d=1e-3, lambda_gp=0.0)
        tr.g_optimizer.write_back_grads = True
        a, x = gan.synthetic_molecules(8, 5, 13, 5, seed=7)
        da, dx = gan.synthetic_molecules(8, 5, 13, 5, seed=8)
        tr.step(da, dx, a, x)
    finally:
        kernels._install_backend_for_tests(None)
    r0, r1 = torch.load(tmp_path / "tr0.pt"), torch.load(tmp_path / "tr1.pt")
    for net, mod in (("G", G), ("D", D)):
        for p, w0, w1 in zip(mod.parameters(), r0[net], r1[net]):
            assert torch.equal(w0, w1)                                            # replicas stay in lock-step
            assert float((w0 - p.detach()).abs().max()) < 5e-5                     # (AdamW's first step is lr * sign-like: bound on the update)
    for p, g0 in zip(G.parameters(), r0["gG"]):
        assert (p.grad is None) == (g0 is None)
        if g0 is not None:
            assert torch.allclose(g0, p.grad, rtol=5e-3, atol=1e-6)
