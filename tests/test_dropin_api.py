"""The reference-facing surface: import paths, constructor signatures, attributes, state-dict keys."""
import inspect

import torch

import druggen_b200 as dg
from conftest import load_golden


def test_src_model_import_paths():
    from src.model.layers import TransformerEncoder, Encoder_Block, MHA, MLP  # noqa: F401
    from src.model.models import Generator, Discriminator, simple_disc  # noqa: F401
    from src.model.loss import discriminator_loss, generator_loss, gradient_penalty  # noqa: F401
    assert TransformerEncoder is dg.TransformerEncoder and Generator is dg.Generator


def test_signatures_match_reference():
    sig = lambda f: list(inspect.signature(f).parameters)  # noqa: E731
    assert sig(dg.TransformerEncoder.__init__)[1:] == ["dim", "depth", "heads", "act", "mlp_ratio", "drop_rate"]
    assert inspect.signature(dg.TransformerEncoder.__init__).parameters["mlp_ratio"].default == 4
    assert inspect.signature(dg.TransformerEncoder.__init__).parameters["drop_rate"].default == 0.1
    assert sig(dg.Encoder_Block.__init__)[1:] == ["dim", "heads", "act", "mlp_ratio", "drop_rate"]
    assert sig(dg.MHA.__init__)[1:] == ["dim", "heads", "attention_dropout"]
    assert sig(dg.MLP.__init__)[1:] == ["in_feat", "hid_feat", "out_feat", "dropout"]
    want = ["act", "vertexes", "edges", "nodes", "dropout", "dim", "depth", "heads", "mlp_ratio"]
    assert sig(dg.Generator.__init__)[1:] == want and sig(dg.Discriminator.__init__)[1:] == want
    assert sig(dg.Generator.forward)[1:] == ["z_e", "z_n"] and sig(dg.TransformerEncoder.forward)[1:] == ["x", "y"]
    from druggen_b200 import gan
    assert sig(gan.discriminator_loss) == ["G", "D", "drug_adj", "drug_annot", "mol_adj", "mol_annot", "batch_size", "device", "lambda_gp"]


def test_state_dict_keys_and_attributes_match_reference_checkpoint():
    g = load_golden("gan_step.npz")
    G = dg.Generator("relu", 9, 5, 13, 0.0, dim=128, depth=1, heads=8, mlp_ratio=3)
    D = dg.Discriminator("relu", 9, 5, 13, 0.0, dim=128, depth=1, heads=8, mlp_ratio=3)
    ref_g = sorted(k[4:] for k in g if k.startswith("wG::"))
    ref_d = sorted(k[4:] for k in g if k.startswith("wD::"))
    assert sorted(G.state_dict()) == ref_g and sorted(D.state_dict()) == ref_d
    for k, v in G.state_dict().items():
        assert tuple(v.shape) == g["wG::" + k].shape, k
    for attr in ("vertexes", "edges", "nodes", "depth", "dim", "heads", "mlp_ratio", "dropout", "features", "transformer_dim"):
        assert hasattr(G, attr) and hasattr(D, attr)
    assert G.features == 9 * 9 * 5 + 9 * 13 and D.node_features == 9 * 128
    for act in ("relu", "leaky", "sigmoid", "tanh"):
        dg.Discriminator(act, 9, 5, 13, 0.0, dim=128, depth=1, heads=8, mlp_ratio=3)


def test_trainer_checkpoint_files_follow_the_reference(tmp_path):
    """GANTrainer.save_model / restore_model (train.py:250-263): the reference's file names, plain state_dicts with the reference's
    keys; restoring lands in the flat AdamW buckets the parameters alias."""
    import torch
    import druggen_b200 as dg
    from druggen_b200 import gan, kernels
    from emul_kernels import EmulBackend
    kernels._install_backend_for_tests(EmulBackend())
    try:
        torch.manual_seed(0)
        mk = lambda cls: cls("relu", 5, 5, 13, 0.0, dim=32, depth=1, heads=4, mlp_ratio=3)  # noqa: E731
        tr = gan.GANTrainer(mk(dg.Generator), mk(dg.Discriminator))
        tr.save_model(str(tmp_path), 2, 9)
        assert sorted(p.name for p in tmp_path.iterdir()) == ["3-10-D.ckpt", "3-10-G.ckpt"]
        want = {k: v.clone() for k, v in tr.G.state_dict().items()}
        assert "TransformerEncoder.Encoder_Blocks.0.attn.out_e.weight" in want and "readout_e.bias" in want
        with torch.no_grad():
            for p in tr.G.parameters():
                p.add_(1.0)
        tr.restore_model(3, 10, str(tmp_path))
        for k, v in tr.G.state_dict().items():
            assert torch.equal(v, want[k]), k
        p0 = next(tr.G.parameters())
        assert p0.data_ptr() == tr.g_optimizer.flat_p.data_ptr()          # still a view of the optimizer's flat buffer
    finally:
        kernels._install_backend_for_tests(None)


def test_block_params_follow_the_state_dict_names():
    """Encoder_Block._params() (written out for speed) is BLOCK_PARAM_NAMES, name by name."""
    import druggen_b200 as dg
    from druggen_b200.block import BLOCK_PARAM_NAMES
    blk = dg.Encoder_Block(128, 8, None, mlp_ratio=3, drop_rate=0.0)
    named = dict(blk.named_parameters())
    assert list(named) != [] and set(named) == set(BLOCK_PARAM_NAMES)
    got = blk._params()
    assert len(got) == len(BLOCK_PARAM_NAMES)
    for nm, t in zip(BLOCK_PARAM_NAMES, got):
        assert t is named[nm], nm
