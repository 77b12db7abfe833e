"""CPU restatement of the reference VQ-VAE model (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Follows /root/reference:
    src/networks/vqvae.py  ResidualLayer :5-26 (in-place ReLU => relu(x) + f(relu(x))), ResidualStack :29-49
                           (ONE weight-tied layer applied n times, then ReLU), Encoder :52-96, Decoder :99-136
    src/models/vqvae.py    VQVAE.__init__ :46-74, forward :76-89, training_step :91-117

``params`` uses the reference state_dict keys; the weight-tied residual layer appears under
``stack.0 .. stack.{n-1}`` exactly like ``[layer] * n`` does in the reference, and only ``stack.0`` is read.
Gradients come from torch autograd on these functional ops (fp32 CPU).

PARITY PINNING: checked against the unmodified reference modules on CPU (tests/test_vqvae.py, where
/root/reference is mounted) and against tests/golden/vqvae_*.npz (tests/golden/make_golden_vqvae.py).
"""
from collections import OrderedDict
from typing import Dict

import torch
import torch.nn.functional as F

from . import vq_oracle


def param_shapes(channels: int, latent_dim: int, num_embeddings: int = 512, h_dim: int = 128, res_h_dim: int = 128,
                 n_res_layers: int = 3) -> "OrderedDict[str, tuple]":
    """name -> shape in the order of reference VQVAE(...).state_dict(): decoder, encoder, quantiser (:64-70)."""
    o = OrderedDict()
    d = "decoder.inverse_conv_stack"
    o[f"{d}.0.weight"] = (latent_dim, h_dim, 3, 3)            # ConvTranspose2d: [in, out, kh, kw]
    o[f"{d}.0.bias"] = (h_dim,)
    for i in range(n_res_layers):
        o[f"{d}.1.stack.{i}.res_block.1.weight"] = (res_h_dim, h_dim, 3, 3)
        o[f"{d}.1.stack.{i}.res_block.3.weight"] = (h_dim, res_h_dim, 1, 1)
    o[f"{d}.2.weight"] = (h_dim, h_dim // 2, 4, 4)
    o[f"{d}.2.bias"] = (h_dim // 2,)
    o[f"{d}.4.weight"] = (h_dim // 2, channels, 4, 4)
    o[f"{d}.4.bias"] = (channels,)
    e = "encoder.conv_stack"
    o[f"{e}.0.weight"] = (latent_dim // 2, channels, 4, 4)
    o[f"{e}.0.bias"] = (latent_dim // 2,)
    o[f"{e}.2.weight"] = (latent_dim, latent_dim // 2, 4, 4)
    o[f"{e}.2.bias"] = (latent_dim,)
    o[f"{e}.4.weight"] = (latent_dim, latent_dim, 3, 3)
    o[f"{e}.4.bias"] = (latent_dim,)
    for i in range(n_res_layers):
        o[f"{e}.5.stack.{i}.res_block.1.weight"] = (res_h_dim, latent_dim, 3, 3)
        o[f"{e}.5.stack.{i}.res_block.3.weight"] = (latent_dim, res_h_dim, 1, 1)
    o["vector_quntizer.embedding"] = (num_embeddings, latent_dim)   # sic
    return o


def init_params(channels, latent_dim, num_embeddings=512, h_dim=128, res_h_dim=128, n_res_layers=3, seed=0,
                codebook_scale=None) -> Dict[str, torch.Tensor]:
    """Seeded weights; tied residual layers share ONE tensor object across their stack.i aliases."""
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for name, shape in param_shapes(channels, latent_dim, num_embeddings, h_dim, res_h_dim, n_res_layers).items():
        if ".stack." in name and ".stack.0." not in name:
            head, tail = name.split(".stack.")
            out[name] = out[f"{head}.stack.0.{tail.split('.', 1)[1]}"]
        elif name.endswith("embedding"):
            s = codebook_scale if codebook_scale is not None else 1.0 / num_embeddings
            out[name] = (torch.rand(shape, generator=g) * 2 - 1) * s
        elif name.endswith("weight"):
            fan_in = shape[1] * shape[2] * shape[3]
            out[name] = (torch.rand(shape, generator=g) * 2 - 1) / fan_in ** 0.5
        else:
            out[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
    return out


def _n_res(p, prefix):
    return len([k for k in p if k.startswith(prefix) and k.endswith("res_block.1.weight")])


def residual_stack(p, prefix, x):
    """networks/vqvae.py:22-26, :45-49."""
    w3, w1 = p[f"{prefix}.stack.0.res_block.1.weight"], p[f"{prefix}.stack.0.res_block.3.weight"]
    for _ in range(_n_res(p, prefix)):
        x = F.relu(x)                      # the in-place ReLU also rewrites the skip input
        x = x + F.conv2d(F.relu(F.conv2d(x, w3, None, 1, 1)), w1)
    return F.relu(x)


def encoder(p, x):
    """networks/vqvae.py:69-96."""
    e = "encoder.conv_stack"
    x = F.relu(F.conv2d(x, p[f"{e}.0.weight"], p[f"{e}.0.bias"], 2, 1))
    x = F.relu(F.conv2d(x, p[f"{e}.2.weight"], p[f"{e}.2.bias"], 2, 1))
    x = F.conv2d(x, p[f"{e}.4.weight"], p[f"{e}.4.bias"], 1, 1)
    return residual_stack(p, f"{e}.5", x)


def decoder(p, z):
    """networks/vqvae.py:114-136."""
    d = "decoder.inverse_conv_stack"
    x = F.conv_transpose2d(z, p[f"{d}.0.weight"], p[f"{d}.0.bias"], 1, 1)
    x = residual_stack(p, f"{d}.1", x)
    x = F.relu(F.conv_transpose2d(x, p[f"{d}.2.weight"], p[f"{d}.2.bias"], 2, 1))
    return F.conv_transpose2d(x, p[f"{d}.4.weight"], p[f"{d}.4.bias"], 2, 1)


def forward(p, imgs, beta=0.25):
    """models/vqvae.py:76-89."""
    z = encoder(p, imgs)
    q, _, _, _ = vq_oracle.vq_forward(z, p["vector_quntizer.embedding"], beta)
    return decoder(p, q).reshape(imgs.shape)


def training_losses(p, imgs, beta=0.25):
    """models/vqvae.py:91-117 -> (total, recon, vq, commit, z_index, encoder_z)."""
    ez = encoder(p, imgs)
    q, vq_loss, commit_loss, idx = vq_oracle.vq_forward(ez, p["vector_quntizer.embedding"], beta)
    dz = ez + (q - ez).detach()
    fake = decoder(p, dz).reshape(imgs.shape)
    recon = F.mse_loss(fake, imgs)
    total = recon + vq_loss + beta * commit_loss
    return total, recon, vq_loss, commit_loss, idx, ez
