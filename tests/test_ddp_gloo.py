"""CPU, world_size 2 over gloo: the data-parallel contract of the training path.

N ranks x (B/N) samples, each seeding its backward with 1/N, followed by ONE all-reduce(SUM) over
the flat gradient arena (igm_b200.ddpm._allreduce) must reproduce the 1-rank gradient of the
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


def _worker(rank, world, port, out_path, buckets=1):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import igm_b200
    from igm_b200.ddpm import _allreduce, _world

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
    unet.ddp_buckets = buckets
    _allreduce(unet, unet._flat_grad)   # the product's exchange: all-reduce over the flat arena (1 or several pieces)
    if rank == 0:
        torch.save(unet._flat_grad.clone(), out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("buckets", [1, 3])
def test_two_rank_gradient_equals_full_batch(tmp_path, buckets):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, _free_port(), out, buckets), nprocs=2, join=True)
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


# ---------------------------------------------------------------------------
# gradient accumulation under data parallelism (the lazy-backward path, ddpm._reduce_into_grad): every backward's
# gradients go to the pending arena, only that delta is all-reduced, then .grad += delta / world.  The C library's two
# entry points on that path (bind, axpy) are replaced by host stand-ins; the exchange logic is the product's own.
# ---------------------------------------------------------------------------
class _HostLib:
    def __init__(self):
        self.binds = []

    def igm_unet_bind_params(self, ctx, params, grads):
        self.binds.append(grads.value)
        return 0

    def igm_grad_axpy(self, ctx, dst, src, alpha, scale, n, stream):
        import ctypes as C

        import numpy as np
        d = np.ctypeslib.as_array((C.c_float * n).from_address(dst.value))
        s_ = np.ctypeslib.as_array((C.c_float * n).from_address(src.value))
        assert alpha is None
        d += np.float32(scale.value) * s_
        return 0


class _HostEngine:
    def __init__(self):
        self.lib, self.ctx, self.grad_target = _HostLib(), None, "grad"

    def check(self, rc):
        assert rc == 0


def _micro(rank, k):
    g = torch.Generator().manual_seed(100 + 10 * rank + k)
    return g


def _accum_worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import igm_b200
    from igm_b200 import ddpm as D
    D._stream = lambda: None
    unet = igm_b200.Unet(dim=32, channels=3, dim_mults=(1, 2))
    unet._engine = eng = _HostEngine()
    unet.attach_grads(zero=True)
    n = unet._flat_grad.numel()
    for k in range(2):                                    # two micro-batches, no zero_grad in between
        fake = torch.randn(n, generator=_micro(rank, k))  # this rank's gradient of micro-batch k

        def run_backward(scale):
            target = unet._pend if eng.grad_target == "pend" else unet._flat_grad
            target.data += scale * fake                   # the kernels ACCUMULATE into the bound arena (no version bump)
        D._reduce_into_grad(unet, eng, run_backward, scalable=True)
    if rank == 0:
        torch.save(unet._flat_grad.clone(), out_path)
    dist.barrier()
    dist.destroy_process_group()
    unet._engine = None


def test_two_rank_gradient_accumulation_reduces_only_the_delta(tmp_path):
    out = str(tmp_path / "acc.pt")
    mp.spawn(_accum_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    n = got.numel()
    want = sum(torch.randn(n, generator=_micro(r, k)) for r in range(2) for k in range(2)) / 2   # sum_k mean_r g[r][k]
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6), (got - want).abs().max()


def _arena_worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from igm_b200.vqvae import ArenaAdam, GradArena
    g = torch.Generator().manual_seed(5)
    ps = [torch.nn.Parameter(torch.randn(4, 3, generator=g)), torch.nn.Parameter(torch.randn(7, generator=g))]
    arena = GradArena(ps)
    opt = ArenaAdam(arena, lr=0.1)
    opt.zero_grad()
    loss = (ps[0] ** 2).sum() * (rank + 1) + (ps[1] * (rank + 2)).sum()
    loss.backward()
    ps[1].grad = ps[1].grad.clone()       # a grad replaced behind the arena's back is copied in again
    arena.all_reduce_mean()
    assert ps[1].grad.data_ptr() == arena.flat.data_ptr() + 4 * arena.offsets[1]
    if rank == 0:
        torch.save([p.grad.clone() for p in ps] + [p.detach().clone() for p in ps], out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_vqvae_grad_arena_all_reduce_is_the_mean(tmp_path):
    out = str(tmp_path / "arena.pt")
    mp.spawn(_arena_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    g0, g1, p0, p1 = torch.load(out)
    assert torch.allclose(g0, 2 * p0 * 1.5) and torch.allclose(g1, torch.full_like(p1, 2.5))
