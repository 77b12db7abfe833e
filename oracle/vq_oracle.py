"""CPU restatement of the reference VQ-VAE quantiser (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Follows /root/reference/src/models/vqvae.py:
    VectorQuantizer.__init__ :14-22   (codebook init U(-1/K, 1/K))
    VectorQuantizer.forward  :24-43   (torch.cdist -> argmin -> gather -> two MSE losses)
    VQVAE.training_step      :91-117  (straight-through estimator :103, loss mix :110)

PARITY PINNING: checked bit-exact against the unmodified reference module on CPU
(tests/test_vq.py, where /root/reference is mounted) and against tests/golden/vq_*.npz
generated from the reference by tests/golden/make_golden_vq.py.
"""
import torch
import torch.nn.functional as F


def init_codebook(num_embeddings: int, latent_dim: int, seed: int = 0) -> torch.Tensor:
    """vqvae.py:16-19 with a seeded generator."""
    g = torch.Generator().manual_seed(seed)
    return torch.zeros(num_embeddings, latent_dim).uniform_(-1 / num_embeddings, 1 / num_embeddings, generator=g)


def vq_forward(z: torch.Tensor, embedding: torch.Tensor, commitment_weight: float):
    """vqvae.py:24-43.  Returns (quant_z [N,C,H,W], vq_loss, commit_loss, z_index [N*H*W])."""
    N, C, H, W = z.shape
    reshape_z = z.reshape(N, C, -1).permute(0, 2, 1).reshape(-1, C)
    dist = torch.cdist(reshape_z, embedding)
    z_index = torch.argmin(dist, dim=1)
    quant_z = embedding[z_index]
    vq_loss = F.mse_loss(reshape_z.detach(), quant_z)
    commit_loss = commitment_weight * F.mse_loss(reshape_z, quant_z.detach())
    quant_z = quant_z.reshape(N, H, W, C).permute(0, 3, 1, 2)
    return quant_z, vq_loss, commit_loss, z_index


def top2_gap(z: torch.Tensor, embedding: torch.Tensor) -> torch.Tensor:
    """Relative gap between the two smallest EXACT (fp64) distances of every latent vector —
    the index-parity criterion of SURVEY.md section 8(c): indices must match where gap > 1e-5."""
    N, C, H, W = z.shape
    r = z.double().reshape(N, C, -1).permute(0, 2, 1).reshape(-1, C)
    d = torch.cdist(r, embedding.double(), compute_mode="donot_use_mm_for_euclid_dist")
    two = d.topk(2, dim=1, largest=False).values
    return (two[:, 1] - two[:, 0]) / two[:, 1].clamp_min(1e-300)
