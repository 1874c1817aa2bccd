"""Drop-in for the reference's ``src/model/layers.py``: same import path, same class names.
``from src.model.layers import TransformerEncoder`` resolves to the B200 implementation."""
from druggen_b200.layers import MLP, MHA, Encoder_Block, TransformerEncoder  # noqa: F401
