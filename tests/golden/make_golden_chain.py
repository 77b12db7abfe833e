"""Golden fixture of a FULL reverse-diffusion chain (T = 1000 steps) from the UNMODIFIED reference, run here on CPU.

    python tests/golden/make_golden_chain.py

Tiny U-Net (CASES["tiny"]: dim 32, mults (1, 2), 16x16, batch 2).  The reference's ``p_sample`` (ddpm.py:390-397) is
executed unmodified for t = 999 ... 0 from a seeded x_T; its per-step draw ``noise_like`` (:268-273) is patched to return
the k-th slice of a seeded noise tensor, which the tests regenerate from the same seed (``chain_inputs``).  Snapshots of
the image after 1, 10, 100 and 1000 steps are kept ("sampler checked after 1, 10, 100 and T steps").
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ddpm_oracle as O  # noqa: E402
from oracle import ref_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASE = (32, 3, (1, 2), 16, 16, 2, 1000)      # = make_golden.CASES["tiny"]
SNAPSHOTS = (1, 10, 100, 1000)


def chain_inputs():
    dim, ch, mults, H, W, B, T = CASE
    g = torch.Generator().manual_seed(31337)
    img = torch.randn(B, ch, H, W, generator=g)
    noise = torch.randn(T, B, ch, H, W, generator=g)
    return img, noise


def main():
    dim, ch, mults, H, W, B, T = CASE
    ref = ref_loader.load("ddpm")
    spec = O.UnetSpec(dim, ch, mults)
    unet = ref.Unet(dim=dim, channels=ch, dim_mults=mults)
    unet.load_state_dict(O.init_params(spec, seed=7))
    gd = ref.GaussianDiffusion(unet, image_size=(H, W), channels=ch, timesteps=T, loss_type="l1")
    img, noise = chain_inputs()
    k = [0]

    def injected(shape, device, repeat=False):
        z = noise[k[0]]
        k[0] += 1
        return z

    orig = ref.noise_like
    ref.noise_like = injected
    out = {}
    try:
        for step, i in enumerate(reversed(range(T)), start=1):
            img = gd.p_sample(img, torch.full((B,), i, dtype=torch.long))
            if step in SNAPSHOTS:
                out[f"after_{step}"] = img.numpy().copy()
    finally:
        ref.noise_like = orig
    np.savez_compressed(os.path.join(HERE, "ddpm_chain_tiny.npz"), **out)
    print({k_: float(np.abs(v).max()) for k_, v in out.items()})


if __name__ == "__main__":
    main()
