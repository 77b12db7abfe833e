"""GPU parity: the CUDA path, called through the C ABI via the host mirror, against the CPU
oracle on the same seeded inputs and against the committed reference fixtures.
Tolerance: 1e-3 relative (rel-L2 and max-abs/max-ref), the bound north_star states for fp32."""
import os

import numpy as np
import pytest
import torch

import igm_b200
from oracle import ddpm_oracle as O
from tests._util import CASES, REL_TOL, assert_close, golden, make_golden, rel_err

pytestmark = pytest.mark.gpu

REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _report(line):
    if os.path.isdir(REPORT):
        with open(os.path.join(REPORT, "parity_report.txt"), "a") as f:
            f.write(line + "\n")


@pytest.fixture(autouse=True, params=[0, 1], ids=["simt", "tcgen05"])
def conv_engine(request, monkeypatch):
    """Every parity test runs on both conv engines (the env var is read at engine creation)."""
    monkeypatch.setenv("IGM_CONV_ENGINE", str(request.param))
    return request.param


def _build(case, training=True):
    dim, ch, mults, H, W, B, T = CASES[case]
    spec = O.UnetSpec(dim, ch, mults)
    params = O.init_params(spec, seed=7)
    unet = igm_b200.Unet(dim=dim, channels=ch, dim_mults=mults)
    unet.load_state_dict(params)
    gd = igm_b200.GaussianDiffusion(unet, image_size=(H, W), channels=ch, timesteps=T, loss_type="l1").cuda()
    return spec, params, gd.denoise_fn, gd


@pytest.mark.parametrize("case", list(CASES))
def test_unet_forward_taps(case):
    spec, params, unet, gd = _build(case)
    x, t, noise, _ = make_golden.inputs(case)
    taps = {}
    with torch.no_grad():
        ref = O.unet_forward(params, spec, x, t, taps=taps)
        out = unet(x.cuda(), t.cuda())
    worst = 0.0
    for name, r in taps.items():
        got = unet.read_tap(name).reshape(r.shape).cpu()
        l2, mx = rel_err(got, r)
        _report(f"{case:12s} tap {name:32s} rel-L2 {l2:.2e} max-rel {mx:.2e}")
        worst = max(worst, l2, mx)
        assert l2 <= REL_TOL and mx <= REL_TOL, f"{case}: tap {name} rel-L2 {l2:.3e} max {mx:.3e}"
    l2, mx = assert_close(out.cpu(), ref, f"{case} unet_out")
    _report(f"{case:12s} unet_out rel-L2 {l2:.2e} max-rel {mx:.2e} (worst tap {worst:.2e})")
    assert_close(out.cpu(), golden(case)["unet_out"], f"{case} unet_out vs reference fixture")
    assert unet.launch_count() > 0


@pytest.mark.parametrize("case", list(CASES))
def test_p_losses_and_gradients(case):
    spec, params, unet, gd = _build(case)
    dim, ch, mults, H, W, B, T = CASES[case]
    x, t, noise, _ = make_golden.inputs(case)
    buf = O.diffusion_buffers(T)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    ref_loss = O.p_losses(p, spec, buf, x, t, noise, "l1")
    ref_grads = torch.autograd.grad(ref_loss, list(p.values()))
    # precondition of an L1 gradient comparison: no residual sits on the kink of |.| (make_golden.SEEDS)
    with torch.no_grad():
        margin = (noise - O.unet_forward(params, spec, O.q_sample(buf, x, t, noise), t)).abs().min().item()
    assert margin > 2e-4, f"{case}: smallest |noise - pred| = {margin:.2e}; pick another input seed"
    loss = gd.p_losses(x.cuda(), t.cuda(), noise.cuda())
    loss.backward()
    g = golden(case)
    assert abs(loss.item() - ref_loss.item()) <= REL_TOL * abs(ref_loss.item())
    assert abs(loss.item() - g["loss_l1"]) <= REL_TOL * abs(g["loss_l1"])
    bad = []
    for (name, prm), rg, gn in zip(unet.named_parameters(), ref_grads, g["grad_norms"]):
        l2, mx = rel_err(prm.grad, rg)
        _report(f"{case:12s} grad {name:44s} rel-L2 {l2:.2e} max-rel {mx:.2e} |g| {rg.norm().item():.3e}")
        if l2 > REL_TOL or mx > REL_TOL:
            bad.append((name, l2, mx))
        assert abs(prm.grad.norm().item() - gn) <= 2 * REL_TOL * gn + 1e-12, f"{name}: |grad| vs reference fixture"
    assert not bad, f"{case}: gradients outside 1e-3: {bad[:5]} ({len(bad)} tensors)"
    # gradients ACCUMULATE like torch autograd: a second backward doubles them
    g1 = unet._flat_grad.clone()
    gd.p_losses(x.cuda(), t.cuda(), noise.cuda()).backward()
    assert_close(unet._flat_grad, 2 * g1, "accumulation", 1e-4)


def test_l2_loss_and_scaled_backward():
    case = "tiny"
    dim, ch, mults, H, W, B, T = CASES[case]
    spec = O.UnetSpec(dim, ch, mults)
    params = O.init_params(spec, seed=7)
    unet = igm_b200.Unet(dim=dim, channels=ch, dim_mults=mults)
    unet.load_state_dict(params)
    gd = igm_b200.GaussianDiffusion(unet, image_size=(H, W), channels=ch, timesteps=T, loss_type="l2").cuda()
    x, t, noise, _ = make_golden.inputs(case)
    buf = O.diffusion_buffers(T)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    ref_loss = O.p_losses(p, spec, buf, x, t, noise, "l2")
    ref_grads = torch.autograd.grad(3.0 * ref_loss, list(p.values()))
    loss = gd.p_losses(x.cuda(), t.cuda(), noise.cuda())
    assert abs(loss.item() - golden(case)["loss_l2"]) <= REL_TOL * golden(case)["loss_l2"]
    (3.0 * loss).backward()
    for (name, prm), rg in zip(gd.denoise_fn.named_parameters(), ref_grads):
        assert_close(prm.grad, rg, f"l2 grad {name}")


def test_unet_autograd_function_dx():
    case = "tiny"
    spec, params, unet, gd = _build(case)
    x, t, noise, _ = make_golden.inputs(case)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    xr = x.clone().requires_grad_(True)
    ref = O.unet_forward(p, spec, xr, t)
    ref_g = torch.autograd.grad(ref, [xr] + list(p.values()), grad_outputs=noise)
    xc = x.cuda().requires_grad_(True)
    out = unet(xc, t.cuda())
    out.backward(noise.cuda())
    assert_close(xc.grad, ref_g[0], "dL/dx")
    for (name, prm), rg in zip(unet.named_parameters(), ref_g[1:]):
        assert_close(prm.grad, rg, f"grad {name}")


@pytest.mark.parametrize("case", list(CASES))
def test_sampler_steps(case):
    spec, params, unet, gd = _build(case, training=False)
    dim, ch, mults, H, W, B, T = CASES[case]
    x, t, noise, step_noise = make_golden.inputs(case)
    buf = O.diffusion_buffers(T)
    g = golden(case)
    for label, t0 in (("hi", T - 1), ("lo", 2)):
        img = gd._run_sampler(noise.clone().cuda(), t0, 3, noise=step_noise.cuda())
        with torch.no_grad():
            ref = O.p_sample_loop(params, spec, buf, noise.clone(), step_noise, t_start=t0, n_steps=3)
        l2, mx = assert_close(img.cpu(), ref, f"{case} sample3_{label}")
        _report(f"{case:12s} sample3_{label} rel-L2 {l2:.2e} max-rel {mx:.2e}")
        assert_close(img.cpu(), g[f"sample3_{label}"], f"{case} sample3_{label} vs reference fixture")
    # q_sample is pure fp32 mul/mul/add: bit-exact
    q = gd.q_sample(x.cuda(), t.cuda(), noise.cuda()).cpu()
    assert torch.equal(q, O.q_sample(buf, x, t, noise))
    assert torch.equal(q, torch.from_numpy(g["q_sample"]))


def test_sampler_graph_replay_equals_stepwise_and_philox_is_seeded():
    case = "tiny"
    spec, params, unet, gd = _build(case, training=False)
    dim, ch, mults, H, W, B, T = CASES[case]
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(B, ch, H, W, generator=gen)
    nz = torch.randn(12, B, ch, H, W, generator=gen)
    a = gd._run_sampler(x.clone().cuda(), 40, 12, noise=nz.cuda())          # 1 eager + 11 graph replays
    b = x.clone().cuda()
    for k in range(12):                                                      # twelve 1-step calls
        b = gd._run_sampler(b, 40 - k, 1, noise=nz[k:k + 1].cuda())
    assert torch.equal(a, b)
    s1 = gd._run_sampler(x.clone().cuda(), 30, 6, seed=11)
    s2 = gd._run_sampler(x.clone().cuda(), 30, 6, seed=11)
    s3 = gd._run_sampler(x.clone().cuda(), 30, 6, seed=12)
    assert torch.equal(s1, s2) and not torch.equal(s1, s3)
    assert torch.isfinite(s1).all()
    # in-kernel normals: mean ~ 0, var ~ 1 (t=0 adds no noise, t>0 does)
    z = gd._run_sampler(torch.zeros(B, ch, H, W).cuda(), 0, 1, seed=3)
    assert torch.isfinite(z).all()


def test_fused_adam_matches_torch_adam():
    case = "tiny"
    spec, params, unet, gd = _build(case)
    x, t, noise, _ = make_golden.inputs(case)
    opt = igm_b200.FusedAdam(unet, lr=1e-3, betas=(0.9, 0.999))
    ref_p = {k: torch.nn.Parameter(v.clone()) for k, v in params.items()}
    ref_opt = torch.optim.Adam(ref_p.values(), lr=1e-3, betas=(0.9, 0.999))
    buf = O.diffusion_buffers(CASES[case][6])
    for step in range(3):
        opt.zero_grad()
        gd.p_losses(x.cuda(), t.cuda(), noise.cuda()).backward()
        opt.step()
        ref_opt.zero_grad()
        O.p_losses(ref_p, spec, buf, x, t, noise, "l1").backward()
        ref_opt.step()
    # compare the UPDATE (p - p0): Adam's first steps are ~sign(g)*lr, so use an lr-scaled tolerance
    n_bad = 0
    for (name, prm), k in zip(unet.named_parameters(), params):
        d = (prm.detach().cpu() - params[k])
        dr = (ref_p[k].detach() - params[k])
        n_bad += int(((d - dr).abs() > 0.05 * 3e-3).sum())
    total = sum(v.numel() for v in params.values())
    assert n_bad <= 1e-3 * total, f"{n_bad}/{total} parameter updates differ"
    # the engine re-packs after the fused step: forward with the new weights matches the oracle
    with torch.no_grad():
        out = unet(x.cuda(), t.cuda()).cpu()
        ref = O.unet_forward({k: v.detach() for k, v in ref_p.items()}, spec, x, t)
    assert_close(out, ref, "forward after 3 Adam steps", 5e-3)


def test_ddpm_lightning_surface():
    from oracle import ref_loader
    torch.manual_seed(0)
    d = igm_b200.DDPM(ref_loader.datamodule_cfg(3, 16, 16), hidden_dim=32, dim_mults=(1, 2), lr=1e-4, b1=0.9,
                      b2=0.999, timesteps=50).cuda()
    opt = d.configure_optimizers()
    imgs = (torch.randn(4, 3, 16, 16) * 0.5).clamp(-1, 1).cuda()
    losses = []
    for i in range(3):
        opt.zero_grad()
        loss = d.training_step((imgs, None), i)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses)) and "train_loss/loss" in d.logged
    d.diffusion_model.sample  # noqa: B018
    out = d.validation_step((imgs, None), 1)
    assert out.fake_image is None and out.others["diffusion"].shape == imgs.shape
    fake = d.diffusion_model.sample(2)
    assert fake.shape == (2, 3, 16, 16) and torch.isfinite(fake).all()


@pytest.mark.parametrize("case,B", [("tiny", 3), ("cifar10", 5)])
def test_odd_batch_loss_gradients_and_sampler(case, B):
    """Batches that do not fill the tile / box granularities (two 8x8 images per MMA tile, image pairs per TMA box,
    per-image pixel tiles of the halo weight-gradient kernel): loss, every gradient and a 2-step sampler chain.
    L2 loss: the L1 gradient sign(pred - noise) flips on rounding-level differences wherever a residual is ~0, which
    moves every gradient by O(1/sqrt(#elements)) and says nothing about the kernels."""
    dim, ch, mults, H, W, _, T = CASES[case]
    spec = O.UnetSpec(dim, ch, mults)
    params = O.init_params(spec, seed=7)
    unet = igm_b200.Unet(dim=dim, channels=ch, dim_mults=mults)
    unet.load_state_dict(params)
    gd = igm_b200.GaussianDiffusion(unet, image_size=(H, W), channels=ch, timesteps=T, loss_type="l2").cuda()
    g = torch.Generator().manual_seed(4321)
    x = (torch.randn(B, ch, H, W, generator=g) * 0.5).clamp(-1, 1)
    t = torch.randint(0, T, (B,), generator=g)
    noise = torch.randn(B, ch, H, W, generator=g)
    step_noise = torch.randn(2, B, ch, H, W, generator=g)
    buf = O.diffusion_buffers(T)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    ref_loss = O.p_losses(p, spec, buf, x, t, noise, "l2")
    ref_grads = torch.autograd.grad(ref_loss, list(p.values()))
    loss = gd.p_losses(x.cuda(), t.cuda(), noise.cuda())
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= REL_TOL * abs(ref_loss.item())
    for (name, prm), rg in zip(gd.denoise_fn.named_parameters(), ref_grads):
        assert_close(prm.grad, rg, f"{case} B={B} grad {name}")
    img = gd._run_sampler(noise.clone().cuda(), 500, 2, noise=step_noise.cuda())
    with torch.no_grad():
        ref = O.p_sample_loop(params, spec, buf, noise.clone(), step_noise, t_start=500, n_steps=2)
    assert_close(img.cpu(), ref, f"{case} B={B} sampler")


def test_eager_training_step_equals_lazy_backward():
    """DDPM.training_step enqueues the backward pass before it reads the loss back and parks the gradients in a pending
    arena; loss.backward() adds d_loss * pending to .grad.  Same gradients as the lazy path, for Lightning's closure order
    (training_step -> zero_grad -> backward), for a scaled loss, and accumulated over two steps; a stale loss raises."""
    from oracle import ref_loader
    torch.manual_seed(0)
    d = igm_b200.DDPM(ref_loader.datamodule_cfg(3, 16, 16), hidden_dim=32, dim_mults=(1, 2), lr=1e-4, b1=0.9,
                      b2=0.999, timesteps=50, loss_type="l2").cuda()
    unet, gd = d.denoising_model, d.diffusion_model
    opt = d.configure_optimizers()
    imgs = (torch.randn(4, 3, 16, 16) * 0.5).clamp(-1, 1).cuda()

    def lazy(scale):
        torch.manual_seed(123)                      # same t / noise draws as the eager call below
        opt.zero_grad()
        loss = gd(imgs)
        (scale * loss).backward()
        return loss.item(), unet._flat_grad.clone()

    def eager(scale):
        torch.manual_seed(123)
        unet._flat_grad.fill_(7.0)                  # stale content that Lightning's zero_grad() clears AFTER training_step
        loss = d.training_step((imgs, None), 0)
        assert unet._engine.grad_target == "pend"
        assert d.logged["train_loss/loss"] == loss.item()     # read back on the copy stream, same 4 bytes
        opt.zero_grad()
        (scale * loss).backward()
        return loss.item(), unet._flat_grad.clone()

    for scale in (1.0, 0.25):
        l0, g0 = lazy(scale)
        l1, g1 = eager(scale)
        assert l0 == l1
        assert_close(g1, g0, f"eager vs lazy gradients, scale {scale}", 1e-5)
    # accumulation over two micro-batches
    opt.zero_grad()
    torch.manual_seed(5); d.training_step((imgs, None), 0).backward()
    torch.manual_seed(6); d.training_step((imgs, None), 1).backward()
    acc = unet._flat_grad.clone()
    opt.zero_grad()
    torch.manual_seed(5); gd(imgs).backward()
    torch.manual_seed(6); gd(imgs).backward()
    assert_close(acc, unet._flat_grad, "eager accumulation", 1e-5)
    # a loss whose pending gradients were overwritten must not silently use the newer ones
    stale = d.training_step((imgs, None), 0)
    d.training_step((imgs, None), 1)
    with pytest.raises(RuntimeError, match="overwritten"):
        stale.backward()


def test_linear_schedule_repeat_noise_and_interpolate():
    """GaussianDiffusion(betas=linear_beta_schedule(T)), p_sample(repeat_noise=True) and interpolate (reference
    ddpm.py:275-279, :268-273 with :390-397, :417-431) against the oracle; torch's CUDA generator is re-seeded to
    reproduce the draws the mirror makes exactly where the reference makes them."""
    case = "tiny"
    dim, ch, mults, H, W, B, _ = CASES[case]
    T = 100
    spec = O.UnetSpec(dim, ch, mults)
    params = O.init_params(spec, seed=7)
    unet = igm_b200.Unet(dim=dim, channels=ch, dim_mults=mults)
    unet.load_state_dict(params)
    gd = igm_b200.GaussianDiffusion(unet, image_size=(H, W), channels=ch, timesteps=T, loss_type="l2",
                                    betas=igm_b200.linear_beta_schedule(T)).cuda()
    buf = O.diffusion_buffers(T, betas=O.linear_beta_schedule(T))
    g = torch.Generator().manual_seed(99)
    x1 = (torch.randn(B, ch, H, W, generator=g) * 0.5).clamp(-1, 1)
    x2 = (torch.randn(B, ch, H, W, generator=g) * 0.5).clamp(-1, 1)
    t = torch.randint(0, T, (B,), generator=g)
    noise = torch.randn(B, ch, H, W, generator=g)
    with torch.no_grad():
        ref_loss = O.p_losses(params, spec, buf, x1, t, noise, "l2")
        loss = gd.p_losses(x1.cuda(), t.cuda(), noise.cuda())
    assert abs(loss.item() - ref_loss.item()) <= REL_TOL * abs(ref_loss.item())
    # p_sample(repeat_noise=True): ONE [1, C, H, W] draw repeated over the batch
    tt = torch.full((B,), 40, dtype=torch.long)
    torch.manual_seed(77)
    out = gd.p_sample(x1.cuda(), tt.cuda(), repeat_noise=True)
    torch.manual_seed(77)
    z = torch.randn((1, ch, H, W), device="cuda").cpu().repeat(B, 1, 1, 1)
    with torch.no_grad():
        ref = O.p_sample(params, spec, buf, x1, tt, z)
    assert_close(out.cpu(), ref, "p_sample(repeat_noise=True)")
    # interpolate at t = 1: q_sample both images at t = 1, mix, one reverse step at t = 0 (which adds no noise)
    torch.manual_seed(78)
    got = gd.interpolate(x1.cuda(), x2.cuda(), t=1, weight=0.3)
    torch.manual_seed(78)
    n1 = torch.randn_like(x1.cuda()).cpu()
    n2 = torch.randn_like(x2.cuda()).cpu()
    t1 = torch.full((B,), 1, dtype=torch.long)
    mix = (1 - 0.3) * O.q_sample(buf, x1, t1, n1) + 0.3 * O.q_sample(buf, x2, t1, n2)
    with torch.no_grad():
        ref = O.p_sample(params, spec, buf, mix, torch.zeros(B, dtype=torch.long), torch.zeros_like(mix))
    assert_close(got.cpu(), ref, "interpolate(t=1)")
    assert got.shape == x1.shape and torch.isfinite(gd.interpolate(x1.cuda(), x2.cuda(), t=5)).all()


def test_full_1000_step_chain_matches_reference_fixture(conv_engine):
    """The M2 product itself: the captured T = 1000 reverse chain with injected noise, checked after 1, 10, 100 and 1000
    steps against the snapshots of the UNMODIFIED reference (tests/golden/make_golden_chain.py) and the oracle.
    Tolerance 1e-3 (rel-L2 and max-abs / max-ref), the north_star bound -- also after 1000 chained U-Net calls."""
    import importlib.util
    spec_ = importlib.util.spec_from_file_location("make_golden_chain", os.path.join(os.path.dirname(__file__), "golden",
                                                                                     "make_golden_chain.py"))
    mc = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(mc)
    dim, ch, mults, H, W, B, T = mc.CASE
    spec = O.UnetSpec(dim, ch, mults)
    params = O.init_params(spec, seed=7)
    unet = igm_b200.Unet(dim=dim, channels=ch, dim_mults=mults)
    unet.load_state_dict(params)
    gd = igm_b200.GaussianDiffusion(unet, image_size=(H, W), channels=ch, timesteps=T).cuda()
    img, noise = mc.chain_inputs()
    fix = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "ddpm_chain_tiny.npz")))
    noise_dev = noise.cuda()
    for n in mc.SNAPSHOTS:
        out = gd._run_sampler(img.clone().cuda(), T - 1, n, noise=noise_dev[:n].contiguous())
        l2, mx = assert_close(out.cpu(), fix[f"after_{n}"], f"chain after {n} steps vs the reference fixture")
        _report(f"chain tiny  after {n:4d} steps rel-L2 {l2:.2e} max-rel {mx:.2e} (engine {conv_engine})")
    # the public call: p_sample_loop draws x_T itself; same chain when the draw and the noise are injected
    torch.manual_seed(11)
    a = gd.p_sample_loop((B, ch, H, W), noise=noise_dev)
    torch.manual_seed(11)
    x_T = torch.randn((B, ch, H, W), device="cuda")
    b = gd._run_sampler(x_T, T - 1, T, noise=noise_dev)
    assert torch.equal(a, b)


def test_backward_of_an_overwritten_forward_raises():
    """The engine keeps ONE forward's activations: summing two losses before one backward, or sampling between a loss and
    its backward, must raise instead of differentiating the wrong activations."""
    spec, params, unet, gd = _build("tiny")
    x, t, noise, step_noise = make_golden.inputs("tiny")
    xc, tc, nc = x.cuda(), t.cuda(), noise.cuda()
    la = gd.p_losses(xc, tc, nc)
    lb = gd.p_losses(xc, tc, nc)
    with pytest.raises(RuntimeError, match="most recent forward"):
        (la + lb).backward()
    lc = gd.p_losses(xc, tc, nc)
    gd.p_sample(xc, torch.full((xc.shape[0],), 5, device="cuda", dtype=torch.long))
    with pytest.raises(RuntimeError, match="most recent forward"):
        lc.backward()
    ua = unet(xc.clone().requires_grad_(True), tc)
    unet(xc, tc)                                   # inference forward in between (no grad needed: same engine buffers)
    ub = unet(xc.clone().requires_grad_(True), tc)
    with pytest.raises(RuntimeError, match="most recent forward"):
        ua.sum().backward()
    ub.sum().backward()                            # the latest one is fine
    ld = gd.p_losses(xc, tc, nc)
    ld.backward()


def test_gradient_buckets_partition_the_arena_and_their_events_order_the_streams(conv_engine):
    """igm_unet_grad_buckets / igm_unet_bucket_wait (the data-parallel exchange's device-side interface): the buckets are
    disjoint ranges that cover the gradient arena, start at parameter-group boundaries in backward-completion order
    (ups | mid | final first, downs.(n-1) .. downs.1, time_mlp + downs.0 last), and a stream that waits for bucket k
    reads that bucket's final gradients while the rest of the backward pass may still be running."""
    import ctypes as C
    spec, params, unet, gd = _build("cifar10")
    x, t, noise, _ = make_golden.inputs("cifar10")
    xc, tc, nc = x.cuda(), t.cuda(), noise.cuda()
    gd.p_losses(xc, tc, nc).backward()
    torch.cuda.synchronize()
    ref = unet._flat_grad.clone()
    e = unet._engine
    lo, hi = (C.c_int64 * 16)(), (C.c_int64 * 16)()
    n = e.lib.igm_unet_grad_buckets(e.ctx, lo, hi, 16)
    assert n == len(unet.dim_mults) + 1
    ranges = [(int(lo[k]), int(hi[k])) for k in range(n)]
    assert ranges[-1][0] == 0 and ranges[0][1] == unet._flat_grad.numel()
    for k in range(1, n):
        assert ranges[k][1] == ranges[k - 1][0]                      # contiguous, walking down the arena
    offs = {name: off for name, off, _ in unet._layout}
    assert ranges[0][0] == offs["ups.0.0.mlp.1.weight"]
    assert ranges[1][0] == offs[f"downs.{len(unet.dim_mults) - 1}.0.mlp.1.weight"]
    assert ranges[-1][1] == offs["downs.1.0.mlp.1.weight"]
    # a second backward: copy every bucket out on a side stream that only waits for that bucket's events
    unet._flat_grad.zero_()
    side = torch.cuda.Stream()
    outs = []
    loss = gd.p_losses(xc, tc, nc)
    loss.backward()
    with torch.cuda.stream(side):
        for k, (a, b) in enumerate(ranges):
            e.check(e.lib.igm_unet_bucket_wait(e.ctx, k, C.c_void_p(side.cuda_stream)))
            outs.append(unet._flat_grad[a:b].clone())
    side.synchronize()
    torch.cuda.synchronize()
    for (a, b), o in zip(ranges, outs):
        assert_close(o, ref[a:b], f"bucket [{a}, {b}) read behind its events", 1e-5)


@pytest.mark.parametrize("case", ["cifar10", "celeba64"])
def test_tensor_core_attention_path_at_small_batch(case, monkeypatch):
    """The tensor-core attention path (out / dq / dv / T as per-image 1x1 convs on the tcgen05 engine, csrc/attention.cu bottom)
    is only taken from 64 K pixels per launch on, i.e. never at the batch sizes of the other parity tests; force it with
    IGM_ATTN_TC_MIN=0 (read at engine creation) and check loss, every gradient and the attention taps against the oracle."""
    monkeypatch.setenv("IGM_CONV_ENGINE", "1")
    monkeypatch.setenv("IGM_ATTN_TC_MIN", "0")
    dim, ch, mults, H, W, B, T = CASES[case]
    spec = O.UnetSpec(dim, ch, mults)
    params = O.init_params(spec, seed=7)
    unet = igm_b200.Unet(dim=dim, channels=ch, dim_mults=mults)
    unet.load_state_dict(params)
    gd = igm_b200.GaussianDiffusion(unet, image_size=(H, W), channels=ch, timesteps=T, loss_type="l2").cuda()
    x, t, noise, _ = make_golden.inputs(case)
    buf = O.diffusion_buffers(T)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    taps = {}
    ref_loss = O.p_losses(p, spec, buf, x, t, noise, "l2")
    ref_grads = torch.autograd.grad(ref_loss, list(p.values()))
    with torch.no_grad():
        O.unet_forward(params, spec, O.q_sample(buf, x, t, noise), t, taps=taps)
    loss = gd.p_losses(x.cuda(), t.cuda(), noise.cuda())
    got_attn = gd.denoise_fn.read_tap("downs.0.2.out").reshape(taps["downs.0.2.out"].shape).cpu()   # training forward: TC path
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= REL_TOL * abs(ref_loss.item())
    assert_close(got_attn, taps["downs.0.2.out"], f"{case}: attention block output through the tensor-core path")
    worst = 0.0
    for (name, prm), rg in zip(gd.denoise_fn.named_parameters(), ref_grads):
        l2, mx = assert_close(prm.grad, rg, f"{case} tensor-core attention: grad {name}")
        worst = max(worst, l2, mx)
    _report(f"{case:12s} tensor-core attention path at B={B}: worst gradient error {worst:.2e}")
