"""One-GPU timing of the CelebA-64 topology (BASELINE.json configs[2]: dims 3-64-128-256-512, 64x64, 32 images per GPU
= global batch 256 over 8 GPUs): train step and denoise step.  Not the headline bench (that is configs[1])."""
import os
import sys
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import igm_b200  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
# optional: topology override "mnist" = reference configs/experiment/ddpm/mnist.yaml (1 channel, 28x28, dim_mults [2, 4])
MNIST = len(sys.argv) > 2 and sys.argv[2] == "mnist"
dev = torch.device("cuda", 0)
torch.manual_seed(0)
S, CHN, MULTS = (28, 1, (2, 4)) if MNIST else (64, 3, (1, 2, 4, 8))
dm = SimpleNamespace(width=S, height=S, channels=CHN, transforms=SimpleNamespace(normalize=True))
model = igm_b200.DDPM(dm, hidden_dim=64, dim_mults=MULTS, timesteps=1000, loss_type="l1", lr=1e-4, b1=0.9, b2=0.999).to(dev)
gd, unet = model.diffusion_model, model.denoising_model
opt = model.configure_optimizers()
x = (torch.randn(B, CHN, S, S, device=dev) * 0.5).clamp(-1, 1)


def step():
    opt.zero_grad()
    loss = gd(x)
    loss.backward()
    opt.step()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
NAME = "mnist28" if MNIST else "celeba64"
print(f"{NAME} B={B}: train step {ms:.3f} ms -> {1e3 / ms:.1f} steps/s/GPU ({B * 1e3 / ms:.0f} img/s), loss {loss.item():.4f}, "
      f"{B * (10.554 if MNIST else 26.195) / ms:.1f} TFLOP/s")
img = torch.randn(B, CHN, S, S, device=dev)
gd._run_sampler(img, 999, 20, seed=1)
torch.cuda.synchronize()
e0.record()
out = gd._run_sampler(img, 999, 100, seed=2)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 100
print(f"{NAME} B={B}: denoise step {ms:.3f} ms -> {B / ms:.2f} samples/s at T=1000, finite={bool(torch.isfinite(out).all())}, "
      f"{B * (3.519 if MNIST else 8.737) / ms:.1f} TFLOP/s")
