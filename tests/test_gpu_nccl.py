"""GPU, 2 ranks over NCCL (needs two visible GPUs; skipped otherwise -- run with `gpurun --gpus 2`): the data-parallel
training path on hardware.  2 ranks x B/2 images must give every rank the gradient of 1 rank x B images (the loss is a
batch mean, no operator mixes samples: SURVEY.md section 8(e)), through BOTH host paths: the lazy backward
(`gd.p_losses(...).backward()`) and the eager backward of `DDPM.training_step`, with gradient accumulation, and the
parameters must be broadcast from rank 0 when the engine is created."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(dim=64, ch=3, mults=(1, 2, 4), H=32, W=32, B=16, T=1000)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    g = torch.Generator().manual_seed(77)
    c = CFG
    x = (torch.randn(c["B"], c["ch"], c["H"], c["W"], generator=g) * 0.5).clamp(-1, 1)
    t = torch.randint(0, c["T"], (c["B"],), generator=g)
    noise = torch.randn(c["B"], c["ch"], c["H"], c["W"], generator=g)
    return x, t, noise


def _model(seed):
    import igm_b200
    from oracle import ddpm_oracle as O
    c = CFG
    spec = O.UnetSpec(c["dim"], c["ch"], c["mults"])
    unet = igm_b200.Unet(dim=c["dim"], channels=c["ch"], dim_mults=c["mults"])
    unet.load_state_dict(O.init_params(spec, seed=seed))
    gd = igm_b200.GaussianDiffusion(unet, image_size=(c["H"], c["W"]), channels=c["ch"], timesteps=c["T"], loss_type="l2").cuda()
    return gd.denoise_fn, gd


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    unet, gd = _model(seed=7 + rank)                  # DIFFERENT weights per rank: the engine must broadcast rank 0's
    x, t, noise = _inputs()
    n = CFG["B"] // world
    sl = slice(rank * n, (rank + 1) * n)
    xs, ts, ns = x[sl].cuda(), t[sl].cuda(), noise[sl].cuda()
    res = {}
    # lazy path, two backwards without zero_grad (accumulation): .grad = 2 x mean over ranks
    unet.attach_grads(zero=True)
    gd.p_losses(xs, ts, ns).backward()
    gd.p_losses(xs, ts, ns).backward()
    res["lazy2"] = unet._flat_grad.clone().cpu()
    res["params"] = unet._flat.clone().cpu()
    # bucketed exchange
    unet.ddp_buckets = 3; unet.ddp_overlap = False
    unet.attach_grads(zero=True)
    gd.p_losses(xs, ts, ns).backward()
    res["lazy_b3"] = unet._flat_grad.clone().cpu()
    unet.ddp_buckets = 1; unet.ddp_overlap = True
    # eager path (what DDPM.training_step uses)
    unet.attach_grads(zero=True)
    gd.eager_backward = True
    loss = gd.p_losses(xs, ts, ns)
    gd.eager_backward = False
    loss.backward()
    res["eager"] = unet._flat_grad.clone().cpu()
    # direct Unet call: parameter gradients are the mean over ranks too
    unet.attach_grads(zero=True)
    out = unet(xs, ts)
    out.backward(ns)
    res["unet"] = unet._flat_grad.clone().cpu()
    torch.save(res, f"{out_path}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_ranks_over_nccl_equal_one_rank_on_the_full_batch(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "r")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    unet, gd = _model(seed=7)
    assert torch.equal(r0["params"], unet._flat.cpu()) and torch.equal(r1["params"], r0["params"]), "rank 0's parameters were not broadcast"
    x, t, noise = _inputs()
    unet.attach_grads(zero=True)
    gd.p_losses(x.cuda(), t.cuda(), noise.cuda()).backward()
    full = unet._flat_grad.clone().cpu()
    unet.attach_grads(zero=True)
    n = CFG["B"] // 2
    o = unet(x.cuda(), t.cuda())
    o.backward(noise.cuda())
    full_unet = unet._flat_grad.clone().cpu()

    def err(a, b):
        return ((a - b).norm() / b.norm()).item()
    for k in r0:
        assert torch.equal(r0[k], r1[k]), f"ranks disagree on {k}"
    e = {"lazy2": err(r0["lazy2"], 2 * full), "lazy_b3": err(r0["lazy_b3"], full), "eager": err(r0["eager"], full),
         # out.backward(noise): each rank differentiates sum over ITS images; the mean over 2 ranks is half the full-batch sum
         "unet": err(r0["unet"], 0.5 * full_unet)}
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        open(os.path.join(d, "nccl_parity.txt"), "a").write(repr(e) + "\n")
    # fp32 summation order differs (split-K over different pixel counts, then the all-reduce): 1e-4, see GRAD_SUM_TOL
    assert all(v <= 2e-4 for v in e.values()), e
