"""B200-native (sm_100a) DDPM U-Net hot path of Victarry/Image-Generation-models.

Host side: thin PyTorch mirror of the reference's class surface
(``src/models/ddpm.py``: Unet, GaussianDiffusion, DDPM) that keeps constructor
signatures, method names and ``state_dict`` keys, and calls the hand-written
CUDA kernels of ``lib/libigm_b200.so`` through its C ABI (include/igm_b200.h).
There is no CPU / eager fallback: importing the engine without the built
library raises.
"""
from . import _lib  # noqa: F401
from .ddpm import (DDPM, FusedAdam, GaussianDiffusion, Unet, ValidationResult, cosine_beta_schedule,  # noqa: F401
                   extract, linear_beta_schedule, noise_like)
from .vqvae import VQVAE, Decoder, Encoder, VectorQuantizer  # noqa: F401
from .pixelcnn import PixelCNN  # noqa: F401
from .callbacks import SampleImagesCallback, get_grid_images  # noqa: F401

__all__ = ["Unet", "GaussianDiffusion", "DDPM", "FusedAdam", "ValidationResult", "VectorQuantizer", "VQVAE", "Encoder", "Decoder", "PixelCNN", "SampleImagesCallback", "get_grid_images",
           "cosine_beta_schedule", "linear_beta_schedule", "extract", "noise_like"]
