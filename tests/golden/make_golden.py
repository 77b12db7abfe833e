"""Generate the golden fixtures of tests/golden/ by RUNNING THE UNMODIFIED REFERENCE
(/root/reference/src/models/ddpm.py) on CPU through oracle/ref_loader.py.

    python tests/golden/make_golden.py

Weights are not stored: they come from oracle.ddpm_oracle.init_params(spec, seed)
(a seeded torch.Generator stream) and are loaded into the reference modules with
load_state_dict, so only inputs' seeds and the reference's outputs are committed.
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

CASES = {
    # name: (dim, channels, dim_mults, H, W, B, timesteps)
    "tiny": (32, 3, (1, 2), 16, 16, 2, 1000),
    "mnist_like": (32, 1, (2, 4), 28, 28, 2, 1000),
    "cifar10": (64, 3, (1, 2, 4), 32, 32, 2, 1000),
    # BASELINE.json configs[2] topology (CelebA 64x64, dim_mults 1-2-4-8, 29.8 M parameters) at batch 1
    "celeba64": (64, 3, (1, 2, 4, 8), 64, 64, 1, 1000),
}


# Input seed per case.  The L1 loss gradient is sign(noise - pred): an element whose residual is at rounding level
# flips its sign between two correct fp32 implementations and moves EVERY gradient by O(1/sqrt(#elements)).
# Seed 1234 leaves a residual of 8.9e-6 in the one-image celeba64 case, so that case uses a seed whose smallest
# |noise - pred| is 6.0e-4 (tiny 7.7e-4, mnist_like 2.8e-4, cifar10 4.5e-4); the GPU test asserts the margin.
SEEDS = {"celeba64": 2000}


def inputs(case):
    dim, ch, mults, H, W, B, T = CASES[case]
    g = torch.Generator().manual_seed(SEEDS.get(case, 1234))
    x = (torch.randn(B, ch, H, W, generator=g) * 0.5).clamp(-1, 1)
    t = torch.randint(0, T, (B,), generator=g)
    noise = torch.randn(B, ch, H, W, generator=g)
    step_noise = torch.randn(3, B, ch, H, W, generator=g)
    return x, t, noise, step_noise


def main():
    ref = ref_loader.load("ddpm")
    only = sys.argv[1:]
    for case, (dim, ch, mults, H, W, B, T) in CASES.items():
        if only and case not in only:
            continue
        spec = O.UnetSpec(dim, ch, mults)
        params = O.init_params(spec, seed=7)
        unet = ref.Unet(dim=dim, channels=ch, dim_mults=mults)
        unet.load_state_dict(params)
        gd = ref.GaussianDiffusion(unet, image_size=(H, W), channels=ch, timesteps=T, loss_type="l1")
        x, t, noise, step_noise = inputs(case)
        out = {}
        with torch.no_grad():
            out["unet_out"] = unet(x, t).numpy()
            out["q_sample"] = gd.q_sample(x, t, noise).numpy()
        loss = gd.p_losses(x, t, noise)
        loss.backward()
        out["loss_l1"] = np.float32(loss.item())
        names = list(dict(unet.named_parameters()).keys())
        out["grad_norms"] = np.array([p.grad.norm().item() for p in unet.parameters()], dtype=np.float64)
        out["grad_heads"] = np.stack([
            np.pad(p.grad.reshape(-1)[:32].numpy(), (0, max(0, 32 - p.numel()))) for p in unet.parameters()])
        gd.loss_type = "l2"
        with torch.no_grad():
            out["loss_l2"] = np.float32(gd.p_losses(x, t, noise).item())
        # three reverse steps from t = T-1 and three ending at t = 0, with injected noise
        k = {"i": 0}
        ref.noise_like = lambda shape, device, repeat=False: step_noise[k["i"]]
        for label, t0 in (("hi", T - 1), ("lo", 2)):
            img = noise.clone()
            with torch.no_grad():
                for j in range(3):
                    k["i"] = j
                    img = gd.p_sample(img, torch.full((B,), t0 - j, dtype=torch.long))
            out[f"sample3_{label}"] = img.numpy()
        np.savez_compressed(os.path.join(HERE, f"ddpm_{case}.npz"), **out)
        print(case, "params", sum(p.numel() for p in unet.parameters()), "loss", out["loss_l1"],
              {k_: getattr(v, "shape", None) for k_, v in out.items()})


if __name__ == "__main__":
    main()
