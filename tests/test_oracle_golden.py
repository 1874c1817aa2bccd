"""Pins the CPU oracle (oracle/encoder_oracle.py) against golden vectors produced by the
unmodified reference modules (oracle/make_golden.py)."""
import torch

from conftest import load_golden, rel_l2, state_from
from oracle import encoder_oracle as orc

TOL = 2e-5  # fp32 CPU vs fp32 CPU, different op order only


def test_encoder_forward_config1():
    g = load_golden("enc_fwd_cfg1.npz")
    p = state_from(g, "w::")
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    xo, yo = orc.encoder_forward(x, y, p, int(g["depth"]), int(g["heads"]))
    assert rel_l2(xo, g["x_out"]) < TOL
    assert rel_l2(yo, g["y_out"]) < TOL


def test_encoder_grads():
    g = load_golden("enc_grad.npz")
    p = {k: v.requires_grad_(True) for k, v in state_from(g, "w::").items()}
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    y = torch.from_numpy(g["y"]).requires_grad_(True)
    xo, yo = orc.encoder_forward(x, y, p, int(g["depth"]), int(g["heads"]))
    ((xo * torch.from_numpy(g["wx"])).sum() + (yo * torch.from_numpy(g["wy"])).sum()).backward()
    assert rel_l2(xo, g["x_out"]) < TOL and rel_l2(yo, g["y_out"]) < TOL
    assert rel_l2(x.grad, g["dx"]) < 1e-4 and rel_l2(y.grad, g["dy"]) < 1e-4
    for k, v in p.items():
        assert rel_l2(v.grad, g["g::" + k]) < 1e-4, k


def _gan(g):
    return orc.OracleGAN(state_from(g, "wG::"), state_from(g, "wD::"), int(g["depth"]), int(g["depth"]),
                         int(g["heads"]), lambda_gp=float(g["lambda_gp"]))


def test_generator_discriminator_forward_and_decode():
    g = load_golden("gan_step.npz")
    gan = _gan(g)
    mol_a, mol_x = torch.from_numpy(g["mol_a"]), torch.from_numpy(g["mol_x"])
    with torch.no_grad():
        node, edge, ns, es = gan.G(mol_a, mol_x)
        d_real = gan.D(torch.from_numpy(g["drug_a"]), torch.from_numpy(g["drug_x"]))
        d_fake = gan.D(es, ns)
    assert rel_l2(node, g["G_node"]) < TOL and rel_l2(edge, g["G_edge"]) < TOL
    assert rel_l2(ns, g["G_node_sample"]) < TOL and rel_l2(es, g["G_edge_sample"]) < TOL
    assert rel_l2(d_real, g["D_real"]) < 1e-4 and rel_l2(d_fake, g["D_fake"]) < 1e-4
    # argmax decode is bit-exact away from ties (inference.py:197-198)
    safe_n = torch.from_numpy(g["node_gap"]) > 1e-4
    safe_e = torch.from_numpy(g["edge_gap"]) > 1e-4
    assert torch.equal(ns.argmax(-1)[safe_n], torch.from_numpy(g["node_argmax"])[safe_n])
    assert torch.equal(es.argmax(-1)[safe_e], torch.from_numpy(g["edge_argmax"])[safe_e])


def test_gradient_penalty_double_backward():
    g = load_golden("gan_step.npz")
    gan = _gan(g)
    t = {k: torch.from_numpy(g[k]) for k in ("drug_a", "drug_x", "G_node_sample", "G_edge_sample",
                                             "eps_edge", "eps_node")}
    gp = orc.gradient_penalty(gan.D, t["drug_x"], t["drug_a"], t["G_node_sample"], t["G_edge_sample"],
                              t["eps_edge"], t["eps_node"])
    gp.backward()
    assert abs(gp.item() - float(g["gp"])) < 1e-4 * max(1.0, abs(float(g["gp"])))
    for k, v in gan.dp_.items():
        assert rel_l2(v.grad, g["gGP_D::" + k]) < 2e-3, k


def test_gan_losses_and_grads():
    g = load_golden("gan_step.npz")
    gan = _gan(g)
    t = {k: torch.from_numpy(g[k]) for k in ("drug_a", "drug_x", "mol_a", "mol_x", "eps_edge", "eps_node")}
    d = gan.d_loss(t["drug_a"], t["drug_x"], t["mol_a"], t["mol_x"], t["eps_edge"], t["eps_node"])
    d.backward()
    assert abs(d.item() - float(g["d_loss"])) < 1e-4 * max(1.0, abs(float(g["d_loss"])))
    for k, v in gan.dp_.items():
        assert rel_l2(v.grad, g["gD_D::" + k]) < 2e-3, k
    gan._zero()
    gl = gan.g_loss(t["mol_a"], t["mol_x"])
    gl.backward()
    assert abs(gl.item() - float(g["g_loss"])) < 1e-4 * max(1.0, abs(float(g["g_loss"])))
    for k, v in gan.gp_.items():
        assert rel_l2(v.grad, g["gG_G::" + k]) < 1e-3, k
    for k, v in gan.dp_.items():
        assert rel_l2(v.grad, g["gG_D::" + k]) < 1e-3, k
