"""Golden fixtures of the VQ quantiser from the UNMODIFIED reference (src/models/vqvae.py:13-43).

    python tests/golden/make_golden_vq.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader, vq_oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {
    # name: (N, D, H, W, K, codebook) ; "init" = the reference's U(-1/K, 1/K) init, "normal" = N(0,1) codes
    "init": (4, 64, 16, 16, 512, "init"),
    "normal": (4, 64, 16, 16, 512, "normal"),
    "small": (3, 32, 7, 5, 40, "normal"),
}


def inputs(case):
    N, D, H, W, K, kind = CASES[case]
    g = torch.Generator().manual_seed(77)
    if kind == "init":
        emb = vq_oracle.init_codebook(K, D, seed=5)
        z = torch.randn(N, D, H, W, generator=g) * (1.0 / K)
    else:
        emb = torch.randn(K, D, generator=g)
        z = torch.randn(N, D, H, W, generator=g)
    return z, emb


def main():
    ref = ref_loader.load("vqvae")
    for case, (N, D, H, W, K, kind) in CASES.items():
        z, emb = inputs(case)
        vq = ref.VectorQuantizer(K, D, 0.25)
        with torch.no_grad():
            vq.embedding.copy_(emb)
        zz = z.clone().requires_grad_(True)
        quant, vq_loss, commit = vq(zz)
        (vq_loss + 0.5 * commit + (quant * quant).sum() * 1e-3).backward()
        idx = torch.cdist(z.reshape(N, D, -1).permute(0, 2, 1).reshape(-1, D), emb).argmin(1)
        np.savez_compressed(os.path.join(HERE, f"vq_{case}.npz"), idx=idx.numpy(), vq_loss=np.float32(vq_loss.item()),
                            commit_loss=np.float32(commit.item()), dz=zz.grad.numpy(),
                            d_emb_norm=np.float64(vq.embedding.grad.norm().item()),
                            d_emb_head=vq.embedding.grad[:8].numpy())
        print(case, "vq_loss", vq_loss.item(), "unique codes", idx.unique().numel())


if __name__ == "__main__":
    main()
