"""CPU: the oracle restatement against (i) the committed golden fixtures produced by the
reference and (ii) the live, unmodified reference when /root/reference is present."""
import numpy as np
import pytest
import torch

from oracle import ddpm_oracle as O
from oracle import ref_loader
from tests._util import CASES, assert_close, golden, make_golden


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_matches_golden(case):
    dim, ch, mults, H, W, B, T = CASES[case]
    spec = O.UnetSpec(dim, ch, mults)
    p = {k: v.requires_grad_(True) for k, v in O.init_params(spec, seed=7).items()}
    buf = O.diffusion_buffers(T)
    x, t, noise, step_noise = make_golden.inputs(case)
    g = golden(case)
    with torch.no_grad():
        assert_close(O.unet_forward(p, spec, x, t), g["unet_out"], "unet_out", 1e-5)
        assert_close(O.q_sample(buf, x, t, noise), g["q_sample"], "q_sample", 1e-6)
        assert abs(O.p_losses(p, spec, buf, x, t, noise, "l2").item() - g["loss_l2"]) < 1e-5
    loss = O.p_losses(p, spec, buf, x, t, noise, "l1")
    assert abs(loss.item() - g["loss_l1"]) < 1e-5
    grads = torch.autograd.grad(loss, list(p.values()))
    norms = np.array([gr.norm().item() for gr in grads])
    np.testing.assert_allclose(norms, g["grad_norms"], rtol=1e-4, atol=1e-9)
    for label, t0 in (("hi", T - 1), ("lo", 2)):
        with torch.no_grad():
            img = O.p_sample_loop(p, spec, buf, noise.clone(), step_noise, t_start=t0, n_steps=3)
        # at t = T-1 the update multiplies eps by ~2e4 before the clamp: allow fp32 noise
        assert_close(img, g[f"sample3_{label}"], f"sample3_{label}", 1e-4)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_oracle_bit_exact_vs_live_reference():
    ref = ref_loader.load("ddpm")
    torch.manual_seed(0)
    d = ref.DDPM(ref_loader.datamodule_cfg(3, 16, 16), hidden_dim=32, dim_mults=(1, 2), lr=1e-4, b1=0.9, b2=0.999)
    spec = O.UnetSpec(32, 3, (1, 2))
    sd = d.denoising_model.state_dict()
    shapes = O.param_shapes(spec)
    assert list(sd.keys()) == list(shapes.keys())
    assert all(tuple(sd[k].shape) == shapes[k] for k in sd)
    buf = O.diffusion_buffers(1000)
    for k in O.SCHEDULE_KEYS:
        assert torch.equal(buf[k], getattr(d.diffusion_model, k)), k
    torch.manual_seed(1)
    x = (torch.randn(3, 3, 16, 16) * 0.5).clamp(-1, 1)
    t = torch.randint(0, 1000, (3,))
    noise = torch.randn(3, 3, 16, 16)
    with torch.no_grad():
        assert torch.equal(d.denoising_model(x, t), O.unet_forward(sd, spec, x, t))
        assert torch.equal(d.diffusion_model.p_losses(x, t, noise), O.p_losses(sd, spec, buf, x, t, noise))
        ref.noise_like = lambda shape, device, repeat=False: noise
        tt = torch.full((3,), 999, dtype=torch.long)
        assert torch.equal(d.diffusion_model.p_sample(x, tt), O.p_sample(sd, spec, buf, x, tt, noise))


def test_oracle_adam_matches_torch():
    torch.manual_seed(0)
    p = {"w": torch.randn(5, 3), "b": torch.randn(3)}
    q = {k: torch.nn.Parameter(v.clone()) for k, v in p.items()}
    opt = torch.optim.Adam(q.values(), lr=1e-2, betas=(0.9, 0.999))
    m = {k: torch.zeros_like(v) for k, v in p.items()}
    v = {k: torch.zeros_like(vv) for k, vv in p.items()}
    for step in range(1, 4):
        g = {k: torch.randn_like(vv) for k, vv in p.items()}
        for k in q:
            q[k].grad = g[k].clone()
        opt.step()
        O.adam_step(p, g, m, v, step, 1e-2, 0.9, 0.999)
    for k in p:
        assert torch.allclose(p[k], q[k].detach(), rtol=1e-6, atol=1e-7)


def test_oracle_full_chain_matches_reference_fixture():
    """The full T = 1000 reverse chain of the oracle against the snapshots the unmodified reference produced
    (tests/golden/make_golden_chain.py): after 1, 10, 100 and 1000 steps."""
    import importlib.util
    import os

    import numpy as np

    from tests._util import GOLDEN
    spec_ = importlib.util.spec_from_file_location("make_golden_chain", os.path.join(GOLDEN, "make_golden_chain.py"))
    mc = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(mc)
    dim, ch, mults, H, W, B, T = mc.CASE
    spec = O.UnetSpec(dim, ch, mults)
    params = O.init_params(spec, seed=7)
    buf = O.diffusion_buffers(T)
    img, noise = mc.chain_inputs()
    fix = dict(np.load(os.path.join(GOLDEN, "ddpm_chain_tiny.npz")))
    done = 0
    for n in mc.SNAPSHOTS:
        img = O.p_sample_loop(params, spec, buf, img, noise[done:n], t_start=T - 1 - done, n_steps=n - done)
        done = n
        ref = torch.from_numpy(fix[f"after_{n}"])
        err = (img - ref).abs().max().item()
        assert err <= 1e-5 * max(ref.abs().max().item(), 1.0), f"after {n} steps: max abs diff {err:.3e}"
