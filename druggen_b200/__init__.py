"""druggen_b200 -- B200-native (sm_100a) graph-transformer encoder hot path of DrugGEN.

Public surface mirrors the reference's ``src/model`` package:
    from druggen_b200.layers import MLP, MHA, Encoder_Block, TransformerEncoder
    from druggen_b200.models import Generator, Discriminator
"""
from . import kernels  # noqa: F401
from .kernels import argmax_last, get_precision, label2onehot, precision, set_precision  # noqa: F401
from .layers import MLP, MHA, Encoder_Block, TransformerEncoder  # noqa: F401
from .models import Discriminator, Generator  # noqa: F401

__all__ = ["MLP", "MHA", "Encoder_Block", "TransformerEncoder", "Generator", "Discriminator",
           "set_precision", "get_precision", "precision", "kernels", "label2onehot", "argmax_last"]
