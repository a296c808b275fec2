"""mfvae_b200 — B200-native (sm_100a) training hot path of the MF-VAE / MAVAE world model.

``model`` and ``trainer`` mirror the reference's ``torch_ver/model.py`` / ``torch_ver/trainer.py``
module surface; the arithmetic runs in ``libmfvae_b200.so`` (C ABI: ``include/mfvae.h``).
"""
from . import _lib  # noqa: F401
from .model import (MAVAE, PackedBatch, Encoder, ActionEncoder, Decoder, reparameterize,  # noqa: F401
                    loss_vae_fn, loss_s_r_vae_fn)
from .trainer import Trainer, FusedAdam, HostStager, create_dataset, cosine_lr  # noqa: F401
from .replay_buffer import DeviceRing, MultiAgentCPPRB, JaxFbxBuffer  # noqa: F401
