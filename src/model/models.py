"""Drop-in for the reference's ``src/model/models.py`` (train.py:20 imports Generator, Discriminator
and simple_disc from here; simple_disc is imported but never instantiated, train.py:203)."""
import torch.nn as nn

from druggen_b200.models import Discriminator, Generator  # noqa: F401


class simple_disc(nn.Module):
    """Flat MLP critic (reference models.py:212-269).  Unused by train.py; kept so the import resolves.
    It has no encoder in it, so it is plain torch."""

    def __init__(self, act, m_dim, vertexes, b_dim):
        super().__init__()
        acts = {"relu": nn.ReLU, "leaky": nn.LeakyReLU, "sigmoid": nn.Sigmoid, "tanh": nn.Tanh}
        if act not in acts:
            raise ValueError("Unsupported activation function: {}".format(act))
        widths = [vertexes * m_dim + vertexes * vertexes * b_dim, 256, 128, 64, 32, 16]
        layers = []
        for fan_in, fan_out in zip(widths[:-1], widths[1:]):
            layers += [nn.Linear(fan_in, fan_out), acts[act]()]
        self.predictor = nn.Sequential(*layers, nn.Linear(16, 1))

    def forward(self, x):
        return self.predictor(x)
