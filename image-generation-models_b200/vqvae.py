"""Host-side mirror of the reference vector quantiser, backed by libigm_b200.so.

``VectorQuantizer(num_embeddings, latent_dim, commitment_weight)`` keeps the constructor, the
``embedding`` parameter (same U(-1/K, 1/K) init, same state_dict key) and the
``forward(z) -> (quant_z, vq_loss, commit_loss)`` contract of reference
``src/models/vqvae.py:13-43``; the distance search, gather, both MSE losses and their gradients
run in the fused CUDA kernels of csrc/vq.cu.  No CPU fallback.
"""
import ctypes as C

import torch
from torch import nn

from . import _lib


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _VQFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, embedding, beta):
        if z.device.type != "cuda":
            raise RuntimeError("libigm_b200 runs on CUDA (B200, sm_100a) only; there is no CPU fallback")
        if z.dtype != torch.float32 or embedding.dtype != torch.float32:
            raise TypeError("libigm_b200 computes in fp32")
        lib = _lib.load()
        z = z.contiguous()
        emb = embedding.contiguous()
        N, D, H, W = z.shape
        K = emb.shape[0]
        idx = torch.empty(N * H * W, dtype=torch.int64, device=z.device)
        quant = torch.empty_like(z)
        losses = torch.empty(2, dtype=torch.float32, device=z.device)
        ws = torch.empty(lib.igm_vq_workspace_floats(N, H * W), dtype=torch.float32, device=z.device)
        rc = lib.igm_vq_forward(_ptr(z), _ptr(emb), _ptr(idx), _ptr(quant), _ptr(losses), N, D, H * W, K, float(beta),
                                _ptr(ws), _stream())
        _lib.check(None, rc)
        ctx.save_for_backward(z, emb, idx)
        ctx.beta = float(beta)
        ctx.mark_non_differentiable(idx)
        return quant, losses[0], losses[1], idx

    @staticmethod
    def backward(ctx, d_quant, d_vq, d_commit, _d_idx):
        z, emb, idx = ctx.saved_tensors
        lib = _lib.load()
        N, D, H, W = z.shape
        dz = torch.empty_like(z) if ctx.needs_input_grad[0] else None
        d_emb = torch.zeros_like(emb) if ctx.needs_input_grad[1] else None
        f = lambda t: None if t is None else t.to(torch.float32).contiguous()
        d_quant, d_vq, d_commit = f(d_quant), f(d_vq), f(d_commit)
        rc = lib.igm_vq_backward(_ptr(z), _ptr(emb), _ptr(idx), _ptr(d_quant), _ptr(d_vq), _ptr(d_commit), ctx.beta,
                                 _ptr(dz), _ptr(d_emb), N, D, H * W, emb.shape[0], _stream())
        _lib.check(None, rc)
        return dz, d_emb, None


class VectorQuantizer(nn.Module):
    """Drop-in for reference ``VectorQuantizer`` (src/models/vqvae.py:13-43)."""

    def __init__(self, num_embeddings, latent_dim, commitment_weight) -> None:
        super().__init__()
        self.embedding = nn.Parameter(
            torch.zeros(num_embeddings, latent_dim).uniform_(-1 / num_embeddings, 1 / num_embeddings))
        self.latent_dim = latent_dim
        self.commitment_weight = commitment_weight

    def forward(self, z):
        quant_z, vq_loss, commit_loss, idx = _VQFn.apply(z, self.embedding, self.commitment_weight)
        self.last_indices = idx   # [N*H*W] int64, the code chosen for every latent vector
        return quant_z, vq_loss, commit_loss
