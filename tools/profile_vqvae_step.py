"""One VQ-VAE training step (BASELINE.json configs[4] per-GPU share) between cudaProfilerStart/Stop, for
`ncu --profile-from-start off`.  Not a benchmark."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench_secondary import build_vqvae  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = build_vqvae(dev, 128)
opt = model.configure_optimizers()
x = (torch.rand(32, 3, 128, 128) * 2 - 1).to(dev)


def step():
    opt.zero_grad()
    loss = model.training_step((x, None), 0)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
