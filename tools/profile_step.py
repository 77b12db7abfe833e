"""One training step (and optionally one sampler step) of the bench workload between
cudaProfilerStart/Stop, for `ncu --profile-from-start off`.  Not a benchmark."""
import argparse
import os
import sys
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import igm_b200  # noqa: E402
from bench import CH, CONFIGS, DIM, T, synth_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cifar10", choices=list(CONFIGS))
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--what", default="train", choices=["train", "sample", "both"])
ap.add_argument("--engine", type=int, default=-1)
args = ap.parse_args()
cfg = CONFIGS[args.config]
H, W, MULTS = cfg["H"], cfg["W"], cfg["mults"]
if args.batch <= 0:
    args.batch = cfg["batch"]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
dm = SimpleNamespace(width=W, height=H, channels=CH, transforms=SimpleNamespace(normalize=True))
model = igm_b200.DDPM(dm, hidden_dim=DIM, dim_mults=MULTS, timesteps=T, loss_type="l1", lr=1e-4, b1=0.9, b2=0.999).to(dev)
gd, unet = model.diffusion_model, model.denoising_model
opt = model.configure_optimizers()
x = synth_batch(args.batch, 0, H, W).to(dev)


def step():
    opt.zero_grad()
    loss = gd(x)
    loss.backward()
    opt.step()


for _ in range(2):
    step()
if args.engine >= 0:
    e = unet._engine
    e.check(e.lib.igm_set_conv_engine(e.ctx, args.engine))
    step()
img = torch.randn(args.batch, CH, H, W, device=dev)
gd._run_sampler(img, T - 1, 1, seed=1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
if args.what in ("train", "both"):
    step()
if args.what in ("sample", "both"):
    gd._run_sampler(img, T - 2, 1, seed=1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
