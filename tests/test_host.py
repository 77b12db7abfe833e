"""CPU: host-side mirror (module tree, state_dict contract, flat arenas) and the C-ABI
library surface (loads, exports every symbol include/igm_b200.h declares).  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import igm_b200
from igm_b200 import _lib
from oracle import ddpm_oracle as O
from oracle import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "igm_b200.h")).read()
    declared = set(re.findall(r"\b(igm_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/igm_b200.h but not exported"
    assert declared == set(_lib.SYMBOLS), "ctypes table and header disagree"
    assert _lib.load().igm_version() >= 100


@pytest.mark.parametrize("cfg", [(64, 3, (1, 2, 4)), (64, 3, (1, 2, 4, 8)), (64, 1, (2, 4)), (32, 3, (1, 2))])
def test_state_dict_contract(cfg):
    dim, ch, mults = cfg
    u = igm_b200.Unet(dim=dim, channels=ch, dim_mults=mults)
    shapes = O.param_shapes(O.UnetSpec(dim, ch, mults))
    sd = u.state_dict()
    assert list(sd.keys()) == list(shapes.keys())
    assert all(tuple(sd[k].shape) == shapes[k] for k in sd)
    # every parameter (and its grad) is a view into the flat arenas at the recorded offset
    for (name, off, shape), p in zip(u._layout, u.parameters()):
        assert p.data_ptr() == u._flat.data_ptr() + 4 * off
        assert p.grad.data_ptr() == u._flat_grad.data_ptr() + 4 * off
        assert off % 4 == 0


def test_flat_arena_tracks_inplace_updates_and_load_state_dict():
    u = igm_b200.Unet(dim=32, channels=3, dim_mults=(1, 2))
    v0 = u._flat._version
    sd = {k: torch.full_like(v, 0.5) for k, v in u.state_dict().items()}
    u.load_state_dict(sd)
    assert u._flat._version != v0
    n = sum(p.numel() for p in u.parameters())
    assert abs(u._flat.sum().item() - 0.5 * n) < 1e-2
    opt = torch.optim.SGD(u.parameters(), lr=1.0)
    for p in u.parameters():
        p.grad.fill_(0.25)
    v1 = u._flat._version
    opt.step()
    assert u._flat._version != v1
    assert abs(u._flat.sum().item() - 0.25 * n) < 1e-2
    opt.zero_grad()          # set_to_none=True drops the views ...
    u.attach_grads()         # ... and attach_grads() restores them onto a zeroed arena
    assert all(p.grad is not None for p in u.parameters())
    assert float(u._flat_grad.abs().sum()) == 0.0


def test_ddpm_module_matches_reference_keys_and_schedule():
    dm = ref_loader.datamodule_cfg(3, 32, 32)
    torch.manual_seed(0)
    d = igm_b200.DDPM(dm, hidden_dim=64, dim_mults=(1, 2, 4), lr=1e-4, b1=0.9, b2=0.999)
    sd = d.state_dict()
    assert len(sd) == 368          # 368 keys for the CIFAR-10 config
    buf = O.diffusion_buffers(1000)
    for k in O.SCHEDULE_KEYS:
        assert torch.equal(sd[f"diffusion_model.{k}"], buf[k]), k
    assert d.hparams.lr == 1e-4 and d.hparams.b1 == 0.9
    if ref_loader.available():
        ref = ref_loader.load("ddpm")
        torch.manual_seed(0)
        r = ref.DDPM(dm, hidden_dim=64, dim_mults=(1, 2, 4), lr=1e-4, b1=0.9, b2=0.999)
        rs = r.state_dict()
        assert list(rs.keys()) == list(sd.keys())
        # same construction order => same default init under the same seed
        assert all(torch.equal(rs[k], sd[k]) for k in rs)


def test_no_cpu_fallback():
    u = igm_b200.Unet(dim=32, channels=3, dim_mults=(1, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        u(torch.zeros(1, 3, 16, 16), torch.zeros(1, dtype=torch.long))


def test_schedule_helpers_and_betas_argument():
    """linear_beta_schedule / cosine_beta_schedule / extract / noise_like (reference ddpm.py:263-291) and the
    ``betas=`` constructor argument (:303-310): mirror == oracle everywhere, == live reference where it is mounted."""
    T = 200
    lin = igm_b200.linear_beta_schedule(T)
    assert lin.dtype == torch.float64 and np.array_equal(lin.numpy(), O.linear_beta_schedule(T))
    assert np.array_equal(igm_b200.cosine_beta_schedule(T), O.cosine_beta_schedule(T))
    u = igm_b200.Unet(dim=32, channels=3, dim_mults=(1, 2))
    gd = igm_b200.GaussianDiffusion(u, image_size=(16, 16), channels=3, timesteps=T, betas=lin)
    buf = O.diffusion_buffers(T, betas=O.linear_beta_schedule(T))
    assert gd.num_timesteps == T
    for k in O.SCHEDULE_KEYS:
        assert torch.equal(getattr(gd, k), buf[k]), k
    a = torch.arange(10.0)
    t = torch.tensor([3, 7])
    assert torch.equal(igm_b200.extract(a, t, (2, 3, 4, 4)), torch.tensor([3.0, 7.0]).reshape(2, 1, 1, 1))
    torch.manual_seed(3)
    z = igm_b200.noise_like((4, 3, 8, 8), "cpu", repeat=True)
    assert z.shape == (4, 3, 8, 8) and torch.equal(z[0], z[3])
    if ref_loader.available():
        ref = ref_loader.load("ddpm")
        assert torch.equal(ref.linear_beta_schedule(T), lin)
        assert np.array_equal(ref.cosine_beta_schedule(T), igm_b200.cosine_beta_schedule(T))
        assert torch.equal(ref.extract(a, t, (2, 3, 4, 4)), igm_b200.extract(a, t, (2, 3, 4, 4)))
        torch.manual_seed(3)
        assert torch.equal(ref.noise_like((4, 3, 8, 8), "cpu", repeat=True), z)
        rgd = ref.GaussianDiffusion(ref.Unet(dim=32, channels=3, dim_mults=(1, 2)), image_size=(16, 16), channels=3,
                                    timesteps=T, betas=lin)
        for k in O.SCHEDULE_KEYS:
            assert torch.equal(getattr(rgd, k), getattr(gd, k)), k


def test_fused_adam_checkpoint_interchanges_with_torch_adam():
    """FusedAdam.state_dict() is torch.optim.Adam's format (per-parameter exp_avg / exp_avg_sq / step), in the order of
    ``denoising_model.parameters()`` the reference hands to Adam (ddpm.py:507): a checkpoint resumes under either."""
    u = igm_b200.Unet(dim=32, channels=3, dim_mults=(1, 2))
    opt = igm_b200.FusedAdam(u, lr=1e-4, betas=(0.9, 0.999))
    g = torch.Generator().manual_seed(0)
    opt._m = torch.randn(u._flat.numel(), generator=g)
    opt._v = torch.rand(u._flat.numel(), generator=g)
    opt._step = 17
    sd = opt.state_dict()
    # -> the reference's optimizer
    ref_params = [torch.nn.Parameter(p.detach().clone()) for p in u.parameters()]
    adam = torch.optim.Adam(ref_params, lr=1e-4, betas=(0.9, 0.999))
    adam.load_state_dict(sd)
    for (name, off, shape), p in zip(u._layout, ref_params):
        st = adam.state[p]
        assert float(st["step"]) == 17
        assert torch.equal(st["exp_avg"].reshape(-1), opt._m[off:off + p.numel()]), name
        assert torch.equal(st["exp_avg_sq"].reshape(-1), opt._v[off:off + p.numel()]), name
    # <- a state the reference's optimizer produced
    for p in ref_params:
        p.grad = torch.randn(p.shape, generator=g)
    adam.step()
    opt2 = igm_b200.FusedAdam(u, lr=1e-4, betas=(0.9, 0.999))
    opt2.load_state_dict(adam.state_dict())
    assert opt2._step == 18
    for (name, off, shape), p in zip(u._layout, ref_params):
        assert torch.equal(opt2._m[off:off + p.numel()], adam.state[p]["exp_avg"].reshape(-1)), name
        assert torch.equal(opt2._v[off:off + p.numel()], adam.state[p]["exp_avg_sq"].reshape(-1)), name
    # a fresh optimizer's (empty) state loads as "no moments yet"
    opt3 = igm_b200.FusedAdam(u)
    opt3.load_state_dict(igm_b200.FusedAdam(u).state_dict())
    assert opt3._m is None and opt3._step == 0
