"""The block-level C-ABI entry points (dg_block_fwd / dg_block_bwd / dg_encoder_fwd) against the same launches
issued one by one from block.py: the forward must be BIT-equal (same kernels, same arguments, same buffers' contents), the
backward equal up to the order of the kernels' atomic reductions (two runs of one path differ the same way)."""
import os

import pytest
import torch

import druggen_b200 as dg
from druggen_b200 import _lib, block
from druggen_b200 import kernels as K
from conftest import rel_l2

pytestmark = pytest.mark.gpu
D, HEADS = 128, 8


def make_block(dev, hid=384, seed=0):
    torch.manual_seed(seed)
    blk = dg.Encoder_Block(D, HEADS, None, mlp_ratio=hid // D, drop_rate=0.0).to(dev)
    with torch.no_grad():
        for nm, p in blk.named_parameters():          # non-trivial LayerNorm affines and biases
            if nm.endswith("bias") or ".ln" in nm or nm.startswith("ln"):
                p.add_(0.1 * torch.randn_like(p))
    return blk


def data(dev, b, n, seed=1):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(b, n, D, generator=g).to(dev), torch.randn(b, n, n, D, generator=g).to(dev)


class native:
    def __init__(self, on):
        self.on = on

    def __enter__(self):
        self.prev = os.environ.get("DRUGGEN_B200_NATIVE_BLOCK")
        os.environ["DRUGGEN_B200_NATIVE_BLOCK"] = "1" if self.on else "0"

    def __exit__(self, *a):
        if self.prev is None:
            os.environ.pop("DRUGGEN_B200_NATIVE_BLOCK", None)
        else:
            os.environ["DRUGGEN_B200_NATIVE_BLOCK"] = self.prev


@pytest.mark.parametrize("b,n", [(3, 9), (2, 45), (40, 45), (5, 90), (1, 4)])
@pytest.mark.parametrize("hid", [384, 128])
def test_block_fwd_bit_equal(cuda_dev, b, n, hid):
    blk = make_block(cuda_dev, hid)
    x, y = data(cuda_dev, b, n)
    params = blk._params()
    with K.precision("bf16"), torch.no_grad():
        assert K.native_block_available(b, n, D, hid)
        l0 = _lib.cuda_backend().launches
        nat = block.block_forward_nograd(x, y, params, HEADS, True, want_stats=True, want_saved=True)
        assert _lib.cuda_backend().launches - l0 == 10          # the ten launches of the forward, issued by the library
        with native(False):
            assert not K.native_block_available(b, n, D, hid)
            ref = block.block_forward_nograd(x, y, params, HEADS, True, want_stats=True, want_saved=True)
    assert torch.equal(nat[0], ref[0]) and torch.equal(nat[1], ref[1])
    for a, r in zip(nat[2], ref[2]):
        assert torch.equal(a.reshape(-1), r.reshape(-1))
    assert set(nat[3]) == set(ref[3])
    for nm in ref[3]:
        assert nat[3][nm].shape == ref[3][nm].shape, nm
        assert torch.equal(nat[3][nm], ref[3][nm]), nm


def test_block_fwd_no_edge_out(cuda_dev):
    blk = make_block(cuda_dev)
    x, y = data(cuda_dev, 4, 45)
    with K.precision("bf16"), torch.no_grad():
        xo, yo = block.block_forward_nograd(x, y, blk._params(), HEADS, False)
        with native(False):
            xr, yr = block.block_forward_nograd(x, y, blk._params(), HEADS, False)
    assert yo is None and yr is None and torch.equal(xo, xr)


@pytest.mark.parametrize("depth,last_edge", [(1, True), (2, True), (3, True), (4, False), (1, False), (2, False), (3, False)])
def test_encoder_fwd_native(cuda_dev, depth, last_edge):
    torch.manual_seed(0)
    enc = dg.TransformerEncoder(dim=D, depth=depth, heads=HEADS, act=None, mlp_ratio=3, drop_rate=0.0).to(cuda_dev)
    enc._discard_final_edge = not last_edge
    x, y = data(cuda_dev, 6, 45)
    with K.precision("bf16"), torch.no_grad():
        xo, yo = enc(x, y)
        with native(False):
            xr, yr = enc(x, y)
    assert torch.equal(xo, xr)
    if last_edge:
        assert torch.equal(yo, yr)
    else:
        assert yo is None and yr is None


def test_encoder_fwd_cuda_graph(cuda_dev):
    """Small encoder forwards replay a captured CUDA graph of dg_encoder_fwd: same bits as the plain launches, new inputs and
    in-place weight updates are honoured, one capture per (shape, weights)."""
    torch.manual_seed(0)
    enc = dg.TransformerEncoder(dim=D, depth=3, heads=HEADS, act=None, mlp_ratio=3, drop_rate=0.0).to(cuda_dev)
    block._GRAPH["cache"].clear()
    assert block._GRAPH["on"]
    with K.precision("bf16"), torch.no_grad():
        for step in range(3):
            x, y = data(cuda_dev, 16, 9, seed=10 + step)
            xo, yo = enc(x, y)
            assert len(block._GRAPH["cache"]) == 1
            block._GRAPH["on"] = False
            try:
                xr, yr = enc(x, y)
            finally:
                block._GRAPH["on"] = True
            assert torch.equal(xo, xr) and torch.equal(yo, yr), step
            for p in enc.parameters():                      # an optimizer step between forwards: the graph re-packs the weights
                p.mul_(1.0 + 0.01 * (step + 1))
        x, y = data(cuda_dev, 8, 9)
        enc(x, y)
        assert len(block._GRAPH["cache"]) == 2              # another shape, another graph


@pytest.mark.parametrize("edge_out,want_params,kept,have_dxo", [
    (True, True, True, True), (True, True, False, True), (True, False, True, True), (True, False, False, True),
    (False, True, False, True), (False, False, False, True), (True, True, True, False)])
@pytest.mark.parametrize("b,n", [(3, 9), (6, 45)])
def test_block_bwd_matches_sequenced(cuda_dev, b, n, edge_out, want_params, kept, have_dxo):
    blk = make_block(cuda_dev)
    params = blk._params()
    x, y = data(cuda_dev, b, n)
    dxo, dyo = data(cuda_dev, b, n, seed=7)
    dxo = dxo if have_dxo else None
    dyo = dyo if edge_out else None

    def run():
        with torch.no_grad():
            xo, yo, stats, saved = block.block_forward_nograd(x, y, params, HEADS, edge_out, want_stats=True, want_saved=kept and edge_out)
            return block.block_backward(x, y, dxo, dyo, params, HEADS, edge_out, want_params, stats, saved if kept else None)

    with K.precision("bf16"):
        dx, dy, gr = run()
        with native(False):
            dxr, dyr, grr = run()
    assert rel_l2(dx, dxr) < 2e-6 and rel_l2(dy, dyr) < 2e-6, (rel_l2(dx, dxr), rel_l2(dy, dyr))
    for nm, g, r in zip(block.BLOCK_PARAM_NAMES, gr, grr):
        assert (g is None) == (r is None), nm
        if g is not None:
            assert g.shape == r.shape and rel_l2(g, r) < 2e-5, (nm, rel_l2(g, r))


@pytest.mark.parametrize("edge_out,kept,have_u", [(True, True, True), (True, False, True), (False, False, True), (True, True, False)])
@pytest.mark.parametrize("b,n", [(3, 9), (6, 45)])
def test_block_bwd_bwd_matches_sequenced(cuda_dev, b, n, edge_out, kept, have_u):
    """dg_block_bwd_bwd (the gradient penalty's second-order pass of a block) against block.py's launch-by-launch sequence."""
    blk = make_block(cuda_dev)
    params = blk._params()
    x, y = data(cuda_dev, b, n)
    dxo, dyo = data(cuda_dev, b, n, seed=7)
    ux, uy = data(cuda_dev, b, n, seed=9)
    dyo = dyo if edge_out else None
    if not have_u:
        uy = None

    def run():
        with torch.no_grad():
            saved = None
            if kept:
                saved = block.block_forward_nograd(x, y, params, HEADS, edge_out, want_stats=True, want_saved=True)[3]
            return block.block_backward_backward(x, y, dxo, dyo, ux, uy, params, HEADS, edge_out, saved)

    with K.precision("bf16"):
        l0 = _lib.cuda_backend().lib.dg_native_launches()
        got = run()
        assert _lib.cuda_backend().lib.dg_native_launches() > l0
        with native(False):
            ref = run()
    bad = {}
    names = ("c_x", "c_y", "c_dxo", "c_dyo") + block.BLOCK_PARAM_NAMES
    for nm, a, r in zip(names, list(got[:4]) + list(got[4]), list(ref[:4]) + list(ref[4])):
        if (a is None) != (r is None):
            bad[nm] = "None mismatch: native %s, sequenced %s" % (a is None, r is None)
        elif a is not None and (a.shape != r.shape or not rel_l2(a, r) < 3e-4):      # (a wrong sequence shows up as 1e-2 .. 1; the order
            # of the fp32 atomic reductions behind c[k] -- cancellation-heavy sums -- moves these outputs by up to ~5e-5 run to run)
            bad[nm] = (tuple(a.shape), tuple(r.shape), rel_l2(a, r))
    assert not bad, bad


def test_probe_times_native_launches(cuda_dev):
    blk = make_block(cuda_dev)
    x, y = data(cuda_dev, 8, 45)
    be = _lib.cuda_backend()
    r = 8 * 45 * 45
    key = f"mlp_fwd[R={r},H=384,fused]"
    with K.precision("bf16"), torch.no_grad():
        be.profile_only = key
        be.profile_reset()
        with native(False):
            block.block_forward_nograd(x, y, blk._params(), HEADS, True)        # one Python-issued launch registers the key's bytes
        for _ in range(3):
            block.block_forward_nograd(x, y, blk._params(), HEADS, True)
        rec = be.profile_summary().get(key)
        be.profile_only = None
    assert rec is not None and rec["n"] == 4 and rec["ms"] > 0.0 and rec["bytes"] == 4 * 2 * r * D * 4, rec


def test_gan_step_native_vs_sequenced(cuda_dev):
    """The whole GAN step (gradient penalty included) with and without the block-level entry points."""
    from druggen_b200 import gan
    res = []
    for on in (True, False):
        torch.manual_seed(0)
        G = dg.Generator("relu", 9, 5, 13, 0.0, dim=D, depth=2, heads=HEADS, mlp_ratio=3).to(cuda_dev)
        Dn = dg.Discriminator("relu", 9, 5, 13, 0.0, dim=D, depth=2, heads=HEADS, mlp_ratio=3).to(cuda_dev)
        tr = gan.GANTrainer(G, Dn)
        a, xx = gan.synthetic_molecules(16, 9, 13, 5, seed=3, device=cuda_dev, labels=True)
        torch.manual_seed(5)
        with K.precision("bf16"), native(on):
            l0 = _lib.cuda_backend().lib.dg_native_launches()
            losses = [tr.step(a, xx, a, xx) for _ in range(2)]
            used = _lib.cuda_backend().lib.dg_native_launches() - l0
        assert (used > 0) == on
        res.append((losses, torch.cat([p.detach().reshape(-1) for p in Dn.parameters()]).clone()))
    (la, pa), (lb, pb) = res
    for (d0, g0), (d1, g1) in zip(la, lb):
        assert abs(d0 - d1) < 1e-4 * max(1.0, abs(d1)) and abs(g0 - g1) < 1e-4 * max(1.0, abs(g1)), (la, lb)
    assert rel_l2(pa, pb) < 1e-6
