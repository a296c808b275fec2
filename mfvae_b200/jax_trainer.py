"""The second front-end's step functions (reference ``jax_ver/trainer.py:42-90``) over the same CUDA engine.

jax_ver differs from torch_ver only in the ELBO weighting -- ``kl_weight = 0.1``, ``r_weight = 0.5`` and
``recons = s_loss * (1 - r_weight) + r_loss * r_weight`` (:42-43,64) -- and in having an explicit jitted
``train_step`` / ``test_step`` pair (:73-90).  The KL term is the same number: ``mean_B(sum over all agents' latents)``
equals torch_ver's sum over agents of batch means.  These functions keep the jax names; the "train state" is the
:class:`mfvae_b200.MAVAE` itself (parameters, Adam moments and step count live in its arena).
"""
from .model import MAVAE, PackedBatch

kl_weight = 0.1     # jax_ver/trainer.py:42
r_weight = 0.5      # jax_ver/trainer.py:43


def loss_weights():
    """(kl_weight, r_weight, s_weight) as jax_ver combines them (trainer.py:64), read at call time."""
    return (kl_weight, r_weight, 1.0 - r_weight)


def train_step(model: MAVAE, batch: PackedBatch, lr: float = 1e-3):
    """jax_ver/trainer.py:73-84 (``optax.adam(1e-3)``, jax_ver/main.py:140): forward, ELBO, gradients, Adam update.
    Returns the device tensor ``[loss, s_loss, r_loss, kl_loss]``."""
    return model.train_step(batch, lr, loss_weights=loss_weights())


def test_step(model: MAVAE, batch: PackedBatch):
    """jax_ver/trainer.py:86-90: forward + ELBO only."""
    return model.test_step(batch, loss_weights=loss_weights())
