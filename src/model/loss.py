"""Drop-in for the reference's ``src/model/loss.py`` (same three functions, same signatures).
The reference's own loss.py also runs unchanged on the druggen_b200 modules; this copy exists
because the reference tree is not present on the GPU box."""
from druggen_b200.gan import discriminator_loss, generator_loss, gradient_penalty  # noqa: F401
