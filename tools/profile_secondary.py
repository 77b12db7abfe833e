"""One PixelCNN sampling call, one VQ lookup and one VQ-VAE training step between cudaProfilerStart/Stop, for
`ncu --profile-from-start off` captures of the secondary paths' kernels.  Not a benchmark."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import igm_b200  # noqa: E402
from bench_secondary import _dm, build_vqvae  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
pcnn = igm_b200.PixelCNN(_dm(1, 28, 28, False), hidden_dim=64).to(dev)
vqvae = build_vqvae(dev, 128)
opt = vqvae.configure_optimizers()
x = (torch.rand(32, 3, 128, 128) * 2 - 1).to(dev)


def vq_step():
    opt.zero_grad()
    loss = vqvae.training_step((x, None), 0)
    loss.backward()
    opt.step()


pcnn.sample((64, 1, 28, 28), seed=1)
for _ in range(2):
    vq_step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
pcnn.sample((64, 1, 28, 28), seed=2)
vq_step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
