"""Probe: is a B=128 sampler chain faster as two concurrent B=64 chains on two streams (latency-bound norm / attention
kernels of one half overlapping the tensor-core convs of the other)?  Two model instances = two engine contexts."""
import os
import sys
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import igm_b200  # noqa: E402
from bench import CH, CONFIGS, DIM, T  # noqa: E402

H, W, MULTS = CONFIGS["cifar10"]["H"], CONFIGS["cifar10"]["W"], CONFIGS["cifar10"]["mults"]

dev = torch.device("cuda", 0)
dm = SimpleNamespace(width=W, height=H, channels=CH, transforms=SimpleNamespace(normalize=True))
NSPLIT = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = 128
STEPS = 100


def make():
    torch.manual_seed(0)
    m = igm_b200.DDPM(dm, hidden_dim=DIM, dim_mults=MULTS, timesteps=T, loss_type="l1", lr=1e-4, b1=0.9, b2=0.999).to(dev)
    return m.diffusion_model


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


one = make()
img = torch.randn(B, CH, H, W, device=dev)
ms1 = timed(lambda: one._run_sampler(img.clone(), T - 1, STEPS, seed=1))
print(f"1 x B={B}: {ms1 / STEPS:.3f} ms per denoise step -> {B / (ms1 / STEPS * T) * 1e3:.1f} samples/s")

gds = [make() for _ in range(NSPLIT)]
streams = [torch.cuda.Stream(device=dev) for _ in range(NSPLIT)]
parts = list(img.chunk(NSPLIT))


def split_run():
    cur = torch.cuda.current_stream(dev)
    for s in streams:
        s.wait_stream(cur)
    for gd, s, p in zip(gds, streams, parts):
        with torch.cuda.stream(s):
            gd._run_sampler(p.clone(), T - 1, STEPS, seed=1)
    for s in streams:
        cur.wait_stream(s)


ms2 = timed(split_run)
print(f"{NSPLIT} x B={B // NSPLIT} concurrent: {ms2 / STEPS:.3f} ms per denoise step -> {B / (ms2 / STEPS * T) * 1e3:.1f} samples/s")
