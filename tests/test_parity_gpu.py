"""Parity of the CUDA path (through the C-ABI) with the CPU oracle and the reference golden vectors.

Tolerances (relative L2 against the fp32 CPU oracle / reference):
  fp32 precision (parity mode) : 1e-3  (north_star bar; measured ~1e-6)
  bf16 precision            : reported, bounded at 5e-2 for outputs (bf16 operand rounding, ~7e-3 expected)
argmax decode: identical wherever the reference top-1/top-2 logit gap exceeds 1e-4.
"""
import pytest
import torch

import druggen_b200 as dg
from conftest import load_golden, rel_l2, state_from
from oracle import encoder_oracle as orc

pytestmark = pytest.mark.gpu
PARITY_TOL = 1e-3
# Parameter gradients are discontinuous at ReLU sign changes (layers.py:52): a forward perturbation eps flips ~eps of the hidden
# units and moves the gradients by ~sqrt(eps).  The fp32 reference evaluated with 8 threads instead of 1 already differs from the
# fp64 truth by 2.2e-3 on the depth-8 G-step gradients (tools/oracle_noise.py, profiles/r02_oracle_noise.json).  The fp32 mode
# (forward error 1.5e-7) happens to stay under 1e-3; the tensor-core parity mode bf16x3 (forward error 2e-6, losses 1e-6) is bounded
# at 5e-3 per gradient tensor -- the reference's own fp32 noise level.
GRAD_TOL = {"fp32": 1e-3, "bf16x3": 5e-3}


def _tc_built():
    try:
        with dg.precision("bf16"):
            a = torch.zeros(128, 128, device="cuda")
            dg.kernels.rows_gemm(a, a, True)
        return True
    except RuntimeError:
        return False


@pytest.fixture(params=["fp32", "bf16x3"])
def parity_mode(request):
    if request.param != "fp32" and not _tc_built():
        pytest.skip("tcgen05 contractions not built")
    with dg.precision(request.param):
        yield request.param


def test_encoder_forward_config1_golden(cuda_dev, parity_mode):
    g = load_golden("enc_fwd_cfg1.npz")
    enc = dg.TransformerEncoder(dim=128, depth=1, heads=8, act=None, mlp_ratio=3, drop_rate=0.0)
    enc.load_state_dict(state_from(g, "w::"))
    enc.to(cuda_dev)
    with torch.no_grad():
        xo, yo = enc(torch.from_numpy(g["x"]).to(cuda_dev), torch.from_numpy(g["y"]).to(cuda_dev))
    assert rel_l2(xo, g["x_out"]) < PARITY_TOL and rel_l2(yo, g["y_out"]) < PARITY_TOL


def test_encoder_config1_b32_vs_oracle(cuda_dev, parity_mode):
    """BASELINE config 1 at its own size: 1 layer, batch 32, N=9, oracle run live on the host."""
    torch.manual_seed(0)
    enc = dg.TransformerEncoder(dim=128, depth=1, heads=8, act=None, mlp_ratio=3, drop_rate=0.0)
    x, y = torch.randn(32, 9, 128), torch.randn(32, 9, 9, 128)
    with torch.no_grad():
        xr, yr = orc.encoder_forward(x, y, dict(enc.state_dict()), 1, 8)
        enc.to(cuda_dev)
        xo, yo = enc(x.to(cuda_dev), y.to(cuda_dev))
    assert rel_l2(xo, xr) < PARITY_TOL and rel_l2(yo, yr) < PARITY_TOL


def test_encoder_grads_golden(cuda_dev, parity_mode):
    g = load_golden("enc_grad.npz")
    enc = dg.TransformerEncoder(dim=128, depth=2, heads=4, act=None, mlp_ratio=3, drop_rate=0.0)
    enc.load_state_dict(state_from(g, "w::"))
    enc.to(cuda_dev)
    x = torch.from_numpy(g["x"]).to(cuda_dev).requires_grad_(True)
    y = torch.from_numpy(g["y"]).to(cuda_dev).requires_grad_(True)
    xo, yo = enc(x, y)
    ((xo * torch.from_numpy(g["wx"]).to(cuda_dev)).sum() + (yo * torch.from_numpy(g["wy"]).to(cuda_dev)).sum()).backward()
    assert rel_l2(xo, g["x_out"]) < PARITY_TOL and rel_l2(yo, g["y_out"]) < PARITY_TOL
    assert rel_l2(x.grad, g["dx"]) < PARITY_TOL and rel_l2(y.grad, g["dy"]) < PARITY_TOL
    for k, v in enc.named_parameters():
        assert rel_l2(v.grad, g["g::" + k]) < PARITY_TOL, k


def test_encoder_n45_depth2_fwd_bwd_vs_oracle(cuda_dev, parity_mode):
    """N=45 (the metric's molecule size), ragged vs the 128-row tiles: 2025 edge rows per molecule."""
    torch.manual_seed(3)
    enc = dg.TransformerEncoder(dim=128, depth=2, heads=8, act=None, mlp_ratio=3, drop_rate=0.0)
    x0, y0 = torch.randn(3, 45, 128), torch.randn(3, 45, 45, 128)
    wx, wy = torch.randn(3, 45, 128), torch.randn(3, 45, 45, 128)
    p = {k: v.detach().clone().requires_grad_(True) for k, v in enc.state_dict().items()}
    x, y = x0.clone().requires_grad_(True), y0.clone().requires_grad_(True)
    xr, yr = orc.encoder_forward(x, y, p, 2, 8)
    ((xr * wx).sum() + (yr * wy).sum()).backward()
    enc.to(cuda_dev)
    xg, yg = x0.to(cuda_dev).requires_grad_(True), y0.to(cuda_dev).requires_grad_(True)
    xo, yo = enc(xg, yg)
    ((xo * wx.to(cuda_dev)).sum() + (yo * wy.to(cuda_dev)).sum()).backward()
    assert rel_l2(xo, xr) < PARITY_TOL and rel_l2(yo, yr) < PARITY_TOL
    assert rel_l2(xg.grad, x.grad) < PARITY_TOL and rel_l2(yg.grad, y.grad) < PARITY_TOL
    for k, v in enc.named_parameters():
        assert rel_l2(v.grad, p[k].grad) < GRAD_TOL[parity_mode], k


def _models(g, dev):
    n, m_dim, b_dim = int(g["n"]), int(g["m_dim"]), int(g["b_dim"])
    G = dg.Generator("relu", n, b_dim, m_dim, 0.0, dim=128, depth=int(g["depth"]), heads=int(g["heads"]), mlp_ratio=3)
    D = dg.Discriminator("relu", n, b_dim, m_dim, 0.0, dim=128, depth=int(g["depth"]), heads=int(g["heads"]), mlp_ratio=3)
    G.load_state_dict(state_from(g, "wG::"))
    D.load_state_dict(state_from(g, "wD::"))
    return G.to(dev), D.to(dev)


def test_gan_step_golden(cuda_dev, parity_mode):
    """train.py:351-384 losses on our CUDA modules vs the reference run: both losses, the gradient
    penalty (double backward through the CUDA kernels), every parameter gradient, argmax decode."""
    g = load_golden("gan_step.npz")
    G, D = _models(g, cuda_dev)
    t = {k: torch.from_numpy(g[k]).to(cuda_dev) for k in ("drug_a", "drug_x", "mol_a", "mol_x", "eps_edge", "eps_node",
                                                           "G_node_sample", "G_edge_sample")}
    with torch.no_grad():
        node, edge, ns, es = G(t["mol_a"], t["mol_x"])
        assert rel_l2(ns, g["G_node_sample"]) < PARITY_TOL and rel_l2(es, g["G_edge_sample"]) < PARITY_TOL
        assert rel_l2(node, g["G_node"]) < PARITY_TOL and rel_l2(edge, g["G_edge"]) < PARITY_TOL
        assert rel_l2(D(t["drug_a"], t["drug_x"]), g["D_real"]) < PARITY_TOL
        safe_n = torch.from_numpy(g["node_gap"]) > 1e-4
        safe_e = torch.from_numpy(g["edge_gap"]) > 1e-4
        assert torch.equal(ns.argmax(-1).cpu()[safe_n], torch.from_numpy(g["node_argmax"])[safe_n])
        assert torch.equal(es.argmax(-1).cpu()[safe_e], torch.from_numpy(g["edge_argmax"])[safe_e])

    gp = orc.gradient_penalty(D, t["drug_x"], t["drug_a"], t["G_node_sample"], t["G_edge_sample"], t["eps_edge"], t["eps_node"])
    gp.backward()
    assert abs(gp.item() - float(g["gp"])) < PARITY_TOL * max(1.0, abs(float(g["gp"])))
    for k, v in D.named_parameters():
        if v.grad is not None:
            assert rel_l2(v.grad, g["gGP_D::" + k]) < 5e-3, k   # second-order: looser, fp32 cancellation in the reference itself
    D.zero_grad(set_to_none=True)

    d_loss = orc.discriminator_loss(G, D, t["drug_a"], t["drug_x"], t["mol_a"], t["mol_x"], t["eps_edge"], t["eps_node"],
                                    float(g["lambda_gp"]))
    d_loss.backward()
    assert abs(d_loss.item() - float(g["d_loss"])) < PARITY_TOL * max(1.0, abs(float(g["d_loss"])))
    for k, v in D.named_parameters():
        if v.grad is None:
            assert float(abs(g["gD_D::" + k]).max()) == 0.0, k
        else:
            assert rel_l2(v.grad, g["gD_D::" + k]) < 5e-3, k
    D.zero_grad(set_to_none=True)
    g_loss = orc.generator_loss(G, D, t["mol_a"], t["mol_x"])
    g_loss.backward()
    assert abs(g_loss.item() - float(g["g_loss"])) < PARITY_TOL * max(1.0, abs(float(g["g_loss"])))
    for k, v in G.named_parameters():
        assert rel_l2(v.grad, g["gG_G::" + k]) < GRAD_TOL[parity_mode], k


def test_bf16_mode_error_is_reported(cuda_dev):
    if not _tc_built():
        pytest.skip("tcgen05 contractions not built")
    g = load_golden("enc_fwd_cfg1.npz")
    enc = dg.TransformerEncoder(dim=128, depth=1, heads=8, act=None, mlp_ratio=3, drop_rate=0.0)
    enc.load_state_dict(state_from(g, "w::"))
    enc.to(cuda_dev)
    with dg.precision("bf16"), torch.no_grad():
        xo, yo = enc(torch.from_numpy(g["x"]).to(cuda_dev), torch.from_numpy(g["y"]).to(cuda_dev))
    ex, ey = rel_l2(xo, g["x_out"]), rel_l2(yo, g["y_out"])
    print(f"bf16 mode rel-L2 vs reference: node {ex:.2e} edge {ey:.2e}")
    assert ex < 5e-2 and ey < 5e-2


def test_roundtrip_properties_full_size(cuda_dev):
    """Size-independent properties at the metric's size (B=64, N=45, 8 layers): permutation
    equivariance over molecules and determinism of the forward; LayerNorm'd outputs have unit
    variance rows."""
    torch.manual_seed(5)
    with dg.precision("fp32"):
        enc = dg.TransformerEncoder(dim=128, depth=8, heads=8, act=None, mlp_ratio=3, drop_rate=0.0).to(cuda_dev)
        x, y = torch.randn(64, 45, 128, device=cuda_dev), torch.randn(64, 45, 45, 128, device=cuda_dev)
        perm = torch.randperm(64, device=cuda_dev)
        with torch.no_grad():
            xo, yo = enc(x, y)
            xo2, yo2 = enc(x, y)
            xp, yp = enc(x[perm].contiguous(), y[perm].contiguous())
        assert torch.equal(xo, xo2) and torch.equal(yo, yo2)
        assert rel_l2(xp, xo[perm]) < 1e-5 and rel_l2(yp, yo[perm]) < 1e-5
        assert abs(float(yo.var(dim=-1, unbiased=False).mean()) - 1.0) < 1e-2


def test_encoder_n90_fwd_bwd_vs_oracle(cuda_dev):
    """N=90 (BASELINE config 5's largest molecule size): 8100 edge rows per molecule, fused and unfused kernels."""
    torch.manual_seed(7)
    enc = dg.TransformerEncoder(dim=128, depth=1, heads=4, act=None, mlp_ratio=3, drop_rate=0.0)
    x0, y0 = torch.randn(2, 90, 128), torch.randn(2, 90, 90, 128)
    wx, wy = torch.randn(2, 90, 128), torch.randn(2, 90, 90, 128)
    p = {k: v.detach().clone().requires_grad_(True) for k, v in enc.state_dict().items()}
    x, y = x0.clone().requires_grad_(True), y0.clone().requires_grad_(True)
    xr, yr = orc.encoder_forward(x, y, p, 1, 4)
    ((xr * wx).sum() + (yr * wy).sum()).backward()
    enc.to(cuda_dev)
    for prec, tol in (("fp32", PARITY_TOL), ("bf16", 5e-2)):
        with dg.precision(prec):
            enc.zero_grad(set_to_none=True)
            xg, yg = x0.to(cuda_dev).requires_grad_(True), y0.to(cuda_dev).requires_grad_(True)
            xo, yo = enc(xg, yg)
            ((xo * wx.to(cuda_dev)).sum() + (yo * wy.to(cuda_dev)).sum()).backward()
            assert rel_l2(xo, xr) < tol and rel_l2(yo, yr) < tol
            assert rel_l2(xg.grad, x.grad) < tol and rel_l2(yg.grad, y.grad) < tol
            for k, v in enc.named_parameters():
                assert rel_l2(v.grad, p[k].grad) < (tol if prec == "fp32" else 0.15), (prec, k)


def test_inference_mode_and_single_molecule(cuda_dev):
    """inference.py:180-198: eval + torch.inference_mode, batch size 1 (its default), argmax decode vs oracle."""
    torch.manual_seed(8)
    G = dg.Generator("relu", 9, 5, 13, 0.0, dim=128, depth=2, heads=8, mlp_ratio=3).eval()
    a, x = orc.synthetic_batch(1, 9, 13, 5, seed=5)
    with torch.no_grad():
        ref = orc.generator_forward(a, x, dict(G.state_dict()), 2, 8)
    G.to(cuda_dev)
    with dg.precision("fp32"), torch.inference_mode():
        node, edge, ns, es = G(a.to(cuda_dev), x.to(cuda_dev))
    assert rel_l2(ns, ref[2]) < PARITY_TOL and rel_l2(es, ref[3]) < PARITY_TOL
    gap_e = ref[3].topk(2, -1).values
    safe = (gap_e[..., 0] - gap_e[..., 1]) > 1e-4
    assert torch.equal(es.argmax(-1).cpu()[safe], ref[3].argmax(-1)[safe])


def test_gan_trainer_step_updates_like_oracle(cuda_dev):
    """One full train.py:351-384 iteration (losses, both backward passes, two AdamW steps) on the CUDA path vs the
    oracle with the same eps draws: losses and the updated weights agree."""
    from druggen_b200 import gan
    torch.manual_seed(9)
    n, bsz = 9, 6
    G = dg.Generator("relu", n, 5, 13, 0.0, dim=128, depth=1, heads=8, mlp_ratio=3)
    D = dg.Discriminator("relu", n, 5, 13, 0.0, dim=128, depth=1, heads=8, mlp_ratio=3)
    ref = orc.OracleGAN(dict(G.state_dict()), dict(D.state_dict()), 1, 1, 8, lr=1e-3)
    a, x = orc.synthetic_batch(bsz, n, 13, 5, seed=3)
    da, dx = orc.synthetic_batch(bsz, n, 13, 5, seed=4)
    G.to(cuda_dev), D.to(cuda_dev)
    with dg.precision("fp32"):
        tr = gan.GANTrainer(G, D, lr_g=1e-3, lr_d=1e-3)
        torch.manual_seed(77)                                   # eps_edge then eps_node are drawn on the device
        d_val, g_val = tr.step(da.to(cuda_dev), dx.to(cuda_dev), a.to(cuda_dev), x.to(cuda_dev))
    torch.manual_seed(77)
    eps_e = torch.rand(bsz, 1, 1, 1, device=cuda_dev).cpu()
    eps_n = torch.rand(bsz, 1, 1, device=cuda_dev).cpu()
    d_ref, g_ref = ref.step(da, dx, a, x, eps_e, eps_n)
    assert abs(d_val - d_ref) < 1e-3 * max(1.0, abs(d_ref)) and abs(g_val - g_ref) < 1e-3 * max(1.0, abs(g_ref))
    for k, v in D.state_dict().items():
        assert rel_l2(v, ref.dp_[k].detach()) < 1e-3, k
    for k, v in G.state_dict().items():
        assert rel_l2(v, ref.gp_[k].detach()) < 1e-3, k


def test_bf16_mode_depth8_error_and_decode_flips_reported(cuda_dev, capsys):
    """Throughput (bf16) mode at the metric's depth: Generator, 8 layers, N=45.  Reports rel-L2 of the logits against
    the fp32 CPU oracle and the fraction of argmax decodes that flip (~7e-3 and 0.3-1 % expected for
    bf16-operand GEMMs); the fp32 mode on the same inputs stays under the 1e-3 parity bar."""
    torch.manual_seed(11)
    G = dg.Generator("relu", 45, 5, 13, 0.0, dim=128, depth=8, heads=8, mlp_ratio=3).eval()
    a, x = orc.synthetic_batch(2, 45, 13, 5, seed=6)
    with torch.no_grad():
        ref = orc.generator_forward(a, x, dict(G.state_dict()), 8, 8)
    G.to(cuda_dev)
    res = {}
    for prec in ("fp32", "bf16"):
        with dg.precision(prec), torch.no_grad():
            _, _, ns, es = G(a.to(cuda_dev), x.to(cuda_dev))
        flips = (es.argmax(-1).cpu() != ref[3].argmax(-1)).float().mean().item()
        res[prec] = (rel_l2(ns, ref[2]), rel_l2(es, ref[3]), flips)
    with capsys.disabled():
        for prec, (en, ee, fl) in res.items():
            print(f"\n[{prec}] depth-8 Generator vs fp32 oracle: node logits rel-L2 {en:.2e}, edge logits rel-L2 {ee:.2e}, "
                  f"edge argmax flips {100 * fl:.2f} %")
    assert res["fp32"][0] < PARITY_TOL and res["fp32"][1] < PARITY_TOL and res["fp32"][2] < 2e-3
    assert res["bf16"][1] < 5e-2 and res["bf16"][2] < 0.05


# measured on B200 (printed by the test below, profiles/r02_parity_depth8.json); the bounds are ~2x the measurement
DEPTH8_BOUNDS = {
    # mode: (loss rel, D-grad rel-L2 over the whole parameter vector, G-grad rel-L2 over the whole parameter vector)
    # fp32: measured 1e-7 / 1.1e-6..3.1e-4 / 5.2e-7 on one box and 8e-8 / 3.1e-4 / 2.19e-3 on another: the second is the SAME single ReLU
    # sign flip the multi-threaded fp32 CPU oracle makes against fp64 (2.192e-3, profiles/r02_oracle_noise.json) -- the atomic
    # flush order of the weight-gradient kernels decides whether a 1e-7 forward perturbation crosses that pre-activation's zero
    "fp32": (1e-3, 5e-3, 5e-3),
    "bf16x3": (1e-3, 5e-3, 5e-3),         # measured 7e-7 / 5.5e-4 / 2.4e-3  (ReLU-flip limited, see GRAD_TOL)
    # bf16 (throughput mode), measured at 8 molecules over four seeds (tools/parity_sweep.py, profiles/r02_parity_sweep.jsonl):
    # loss 1.6e-4..2.2e-3, D grads 2.9e-2..8.2e-2, G grads 2.7e-2..7.5e-2.  At 2 molecules the same mode ranges 3.7e-2..3.1e-1 from
    # seed to seed -- and moves by as much when ANY input is perturbed in its last fp32 bit (a Discriminator-head ReLU unit a
    # rounding error away from zero carries a macroscopic share of a 2-molecule gradient) -- so this mode is bounded at batch 8
    "bf16": (1e-2, 0.16, 0.16),
}
DEPTH8_BATCH = {"fp32": 2, "bf16x3": 2, "bf16": 8}


def _flat(grads):
    return torch.cat([g.double().flatten().cpu() for g in grads])


@pytest.mark.parametrize("mode", ["fp32", "bf16x3", "bf16"])
def test_gan_step_depth8_n45_vs_oracle(cuda_dev, mode, capsys):
    """The step the bench times -- depth 8, N = 45, D loss with the gradient penalty's double backward, G loss -- in EVERY
    precision mode against the CPU oracle evaluated in fp64: both losses, every Discriminator gradient of the D-step (gradient penalty
    included), every Generator gradient of the G-step.  Prints the errors (and stores them for profiles/) and bounds them."""
    import json
    import os
    if mode not in dg.kernels.PRECISIONS:
        pytest.skip(f"precision {mode} not built")
    if mode != "fp32" and not _tc_built():
        pytest.skip("tcgen05 contractions not built")
    torch.manual_seed(21)
    n, bsz = 45, DEPTH8_BATCH[mode]
    G = dg.Generator("relu", n, 5, 13, 0.0, dim=128, depth=8, heads=8, mlp_ratio=3)
    D = dg.Discriminator("relu", n, 5, 13, 0.0, dim=128, depth=8, heads=8, mlp_ratio=3)
    # the oracle in fp64: the fp32 oracle's own gradients carry up to 2e-3 of thread-count-dependent noise (GRAD_TOL above)
    f64 = lambda t: t.double()  # noqa: E731
    ref = orc.OracleGAN({k: f64(v) for k, v in G.state_dict().items()}, {k: f64(v) for k, v in D.state_dict().items()}, 8, 8, 8)
    a, x = orc.synthetic_batch(bsz, n, 13, 5, seed=3)
    da, dx = orc.synthetic_batch(bsz, n, 13, 5, seed=4)
    eps_e, eps_n = torch.rand(bsz, 1, 1, 1), torch.rand(bsz, 1, 1)
    d_ref = ref.d_loss(f64(da), f64(dx), f64(a), f64(x), f64(eps_e), f64(eps_n))
    d_ref.backward()
    gD_ref = {k: v.grad.clone() for k, v in ref.dp_.items() if v.grad is not None and float(v.grad.abs().max()) > 1e-12}
    ref._zero()
    g_ref = ref.g_loss(f64(a), f64(x))
    g_ref.backward()
    gG_ref = {k: v.grad.clone() for k, v in ref.gp_.items()}
    G.to(cuda_dev), D.to(cuda_dev)
    dev = lambda t: t.to(cuda_dev)  # noqa: E731
    with dg.precision(mode):
        d = orc.discriminator_loss(G, D, dev(da), dev(dx), dev(a), dev(x), dev(eps_e), dev(eps_n), 10.0)
        d.backward()
        gD = {k: v.grad.clone() for k, v in D.named_parameters() if v.grad is not None}
        G.zero_grad(set_to_none=True), D.zero_grad(set_to_none=True)
        g = orc.generator_loss(G, D, dev(a), dev(x))
        g.backward()
        gG = {k: v.grad.clone() for k, v in G.named_parameters()}
    assert set(gD_ref) <= set(gD)       # (node_mlp.6.bias: exactly 0 in the oracle -- fake - real + gp -- and rounding noise here)
    rec = {"mode": mode, "depth": 8, "atoms": n, "batch": bsz,
           "d_loss_rel": abs(d.item() - d_ref.item()) / max(1.0, abs(d_ref.item())),
           "g_loss_rel": abs(g.item() - g_ref.item()) / max(1.0, abs(g_ref.item())),
           "D_grads_rel_l2_all": rel_l2(_flat(gD[k] for k in gD_ref), _flat(gD_ref.values())),
           "G_grads_rel_l2_all": rel_l2(_flat(gG[k] for k in gG_ref), _flat(gG_ref.values())),
           "D_grads_rel_l2_worst_tensor": max(rel_l2(gD[k], v) for k, v in gD_ref.items()),
           "G_grads_rel_l2_worst_tensor": max(rel_l2(gG[k], v) for k, v in gG_ref.items())}
    with capsys.disabled():
        print("\n[parity depth8] " + json.dumps(rec))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "parity_depth8.jsonl"), "a") as f:
            f.write(json.dumps(rec) + "\n")
    b_loss, b_d, b_g = DEPTH8_BOUNDS[mode]
    assert rec["d_loss_rel"] < b_loss and rec["g_loss_rel"] < b_loss, rec
    assert rec["D_grads_rel_l2_all"] < b_d and rec["G_grads_rel_l2_all"] < b_g, rec
