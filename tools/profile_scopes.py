"""Warm-cache per-launch-scope timings of one training step (CUDA events around every launch scope, via
IGM_PROFILE_DUMP).  Complements the ncu launch list, whose per-kernel times are cold-cache.  Not a benchmark."""
import os
import sys
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/scopes.txt"
if os.path.exists(out):
    os.remove(out)
os.environ["IGM_PROFILE_DUMP"] = out
import torch  # noqa: E402

import igm_b200  # noqa: E402
from bench import CH, CONFIGS, DIM, T, synth_batch as _synth  # noqa: E402

H, W, MULTS = CONFIGS["cifar10"]["H"], CONFIGS["cifar10"]["W"], CONFIGS["cifar10"]["mults"]


def synth_batch(B, seed):
    return _synth(B, seed, H, W)


dev = torch.device("cuda", 0)
torch.manual_seed(0)
dm = SimpleNamespace(width=W, height=H, channels=CH, transforms=SimpleNamespace(normalize=True))
model = igm_b200.DDPM(dm, hidden_dim=DIM, dim_mults=MULTS, timesteps=T, loss_type="l1", lr=1e-4, b1=0.9, b2=0.999).to(dev)
gd, unet = model.diffusion_model, model.denoising_model
opt = model.configure_optimizers()
x = synth_batch(128, 0).to(dev)


def step():
    opt.zero_grad()
    loss = gd(x)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
unet.profile_start()
step()
prof = unet.profile_stop()
tot = sum(v["ms"] for v in prof.values())
print("scopes total %.3f ms" % tot, {k: round(v["ms"], 3) for k, v in prof.items() if v["launches"]})
