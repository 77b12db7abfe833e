"""CPU, world_size 2 over gloo: the data-parallel contract of the training path.

N ranks x (B/N) samples, each seeding its backward with 1/N, followed by ONE all-reduce(SUM) over
the flat gradient arena (igm_b200.ddpm._allreduce_grads) must reproduce the 1-rank gradient of the
full batch -- exactly what the GPU path does with NCCL.  The per-rank gradient is produced by the
oracle here (no GPU in this container); the host-side exchange code is the product's own.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ddpm_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import igm_b200
    from igm_b200.ddpm import _allreduce_grads, _world

    dim, ch, mults, H, W, B, T = 32, 3, (1, 2), 8, 8, 4, 100
    spec = O.UnetSpec(dim, ch, mults)
    params = O.init_params(spec, seed=3)
    buf = O.diffusion_buffers(T)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, ch, H, W, generator=g)
    t = torch.randint(0, T, (B,), generator=g)
    noise = torch.randn(B, ch, H, W, generator=g)

    unet = igm_b200.Unet(dim=dim, channels=ch, dim_mults=mults)
    unet.load_state_dict(params)
    assert _world() == world
    # this rank's shard of the batch; backward seeded with 1/world like _PLossesFn.backward
    sl = slice(rank * B // world, (rank + 1) * B // world)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    loss = O.p_losses(p, spec, buf, x[sl], t[sl], noise[sl], "l1")
    grads = torch.autograd.grad(loss * (1.0 / world), list(p.values()))
    unet.attach_grads(zero=True)
    for prm, gr in zip(unet.parameters(), grads):
        prm.grad.copy_(gr)
    _allreduce_grads(unet)          # the product's exchange: one all-reduce over the flat arena
    if rank == 0:
        torch.save(unet._flat_grad.clone(), out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_equals_full_batch(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    import igm_b200
    dim, ch, mults, H, W, B, T = 32, 3, (1, 2), 8, 8, 4, 100
    spec = O.UnetSpec(dim, ch, mults)
    params = O.init_params(spec, seed=3)
    buf = O.diffusion_buffers(T)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, ch, H, W, generator=g)
    t = torch.randint(0, T, (B,), generator=g)
    noise = torch.randn(B, ch, H, W, generator=g)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    full = torch.autograd.grad(O.p_losses(p, spec, buf, x, t, noise, "l1"), list(p.values()))
    unet = igm_b200.Unet(dim=dim, channels=ch, dim_mults=mults)
    ref = torch.zeros_like(unet._flat_grad)
    for (name, off, shape), gr in zip(unet._layout, full):
        ref[off:off + gr.numel()] = gr.reshape(-1)
    err = (got - ref).norm() / ref.norm()
    assert err < 1e-5, f"2-rank mean gradient differs from the full-batch gradient: {err:.2e}"
