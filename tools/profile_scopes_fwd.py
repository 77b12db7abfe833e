"""Warm-cache per-launch-scope timings of one inference forward (the sampler's U-Net call) at B=128."""
import os
import sys
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/scopes_fwd.txt"
if os.path.exists(out):
    os.remove(out)
os.environ["IGM_PROFILE_DUMP"] = out
import torch  # noqa: E402

import igm_b200  # noqa: E402
from bench import CH, CONFIGS, DIM, T  # noqa: E402

H, W, MULTS = CONFIGS["cifar10"]["H"], CONFIGS["cifar10"]["W"], CONFIGS["cifar10"]["mults"]

dev = torch.device("cuda", 0)
torch.manual_seed(0)
dm = SimpleNamespace(width=W, height=H, channels=CH, transforms=SimpleNamespace(normalize=True))
model = igm_b200.DDPM(dm, hidden_dim=DIM, dim_mults=MULTS, timesteps=T, loss_type="l1", lr=1e-4, b1=0.9, b2=0.999).to(dev)
gd, unet = model.diffusion_model, model.denoising_model
img = torch.randn(128, CH, H, W, device=dev)
gd._run_sampler(img, T - 1, 3, seed=1)
torch.cuda.synchronize()
unet.profile_start()
gd._run_sampler(img, T - 1, 1, seed=1)    # one eager step (no graph replay with n_steps = 1)
prof = unet.profile_stop()
print("fwd scopes total %.3f ms" % sum(v["ms"] for v in prof.values()), {k: round(v["ms"], 3) for k, v in prof.items() if v["launches"]})
