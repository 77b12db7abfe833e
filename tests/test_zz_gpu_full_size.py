"""GPU: BASELINE.json configs[1] (CIFAR-10 U-Net, batch 128 on one GPU) and configs[2] (CelebA-64 U-Net, 256 / 8 = 32
images per GPU) at their FULL per-GPU sizes, where the CPU oracle cannot process the whole batch in seconds.  Every operator of the path is per-sample (GroupNorm / LayerNorm statistics, the
linear-attention softmax and context are all computed inside one image; there is no BatchNorm: SURVEY.md section 8(e)),
which gives size-independent properties:

  * the rows of a batch-128 result equal the oracle run on those samples alone (parity proper, on a subset);
  * the rows equal the CUDA path run on a small batch of the same samples (batch independence), and a repeat run of the
    inference forward reproduces the result;
  * the gradient of the batch-mean loss over 128 samples is the mean of the gradients over its four quarters,
    and the loss the mean of their losses ("a checksum of checksums").

(The file name sorts last on purpose: these are the heaviest tests of the suite.)"""
import pytest
import torch

import igm_b200
from oracle import ddpm_oracle as O
from tests._util import REL_TOL, assert_close

pytestmark = pytest.mark.gpu

DIM, CH, T = 64, 3, 1000
CONFIGS = {
    # name: (dim_mults, H, W, batch per GPU)
    "cifar10_b128": ((1, 2, 4), 32, 32, 128),
    "celeba64_b32": ((1, 2, 4, 8), 64, 64, 32),
}


def _subset(B):
    return [0, B // 2 - 1, B - 1]


def _build(cfg, loss_type="l1"):
    mults, H, W, B = CONFIGS[cfg]
    spec = O.UnetSpec(DIM, CH, mults)
    params = O.init_params(spec, seed=7)
    unet = igm_b200.Unet(dim=DIM, channels=CH, dim_mults=mults)
    unet.load_state_dict(params)
    gd = igm_b200.GaussianDiffusion(unet, image_size=(H, W), channels=CH, timesteps=T, loss_type=loss_type).cuda()
    return spec, params, gd.denoise_fn, gd


def _inputs(cfg, seed=2024):
    mults, H, W, B = CONFIGS[cfg]
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(B, CH, H, W, generator=g) * 0.5).clamp(-1, 1)
    t = torch.randint(0, T, (B,), generator=g)
    noise = torch.randn(B, CH, H, W, generator=g)
    return x, t, noise


@pytest.mark.parametrize("cfg", list(CONFIGS))
def test_full_batch_forward_rows_match_oracle_and_small_batches(cfg):
    spec, params, unet, gd = _build(cfg)
    x, t, _ = _inputs(cfg)
    SUBSET = _subset(x.shape[0])
    with torch.no_grad():
        full = unet(x.cuda(), t.cuda())
        again = unet(x.cuda(), t.cuda())
        assert_close(again, full, "repeat run of the inference forward", 1e-6)
        assert torch.isfinite(full).all()
        ref = O.unet_forward(params, spec, x[SUBSET], t[SUBSET])             # three samples on the CPU oracle
        small = unet(x[SUBSET].cuda(), t[SUBSET].cuda())                     # the same three as a batch of their own
    assert_close(full[SUBSET].cpu(), ref, f"{cfg}: rows of the full-batch forward vs the oracle")
    assert_close(small.cpu(), full[SUBSET].cpu(), "batch independence of the forward", 1e-4)


@pytest.mark.parametrize("cfg", list(CONFIGS))
def test_full_batch_sampler_rows_match_oracle(cfg):
    spec, params, unet, gd = _build(cfg)
    mults, H, W, B = CONFIGS[cfg]
    SUBSET = _subset(B)
    g = torch.Generator().manual_seed(7)
    img = torch.randn(B, CH, H, W, generator=g)
    step_noise = torch.randn(2, B, CH, H, W, generator=g)
    buf = O.diffusion_buffers(T)
    out = gd._run_sampler(img.clone().cuda(), T - 1, 2, noise=step_noise.cuda())
    with torch.no_grad():
        ref = O.p_sample_loop(params, spec, buf, img[SUBSET].clone(), step_noise[:, SUBSET].contiguous(), t_start=T - 1, n_steps=2)
    assert_close(out[SUBSET].cpu(), ref, f"{cfg}: rows of two full-batch denoise steps vs the oracle")


@pytest.mark.parametrize("cfg", list(CONFIGS))
def test_full_batch_gradient_is_the_mean_of_its_quarters(cfg):
    # L2 loss: the L1 sign gradient is discontinuous, the property is about summation, not about the kink
    spec, params, unet, gd = _build(cfg, "l2")
    x, t, noise = _inputs(cfg, seed=99)
    B = x.shape[0]
    SUBSET = _subset(B)
    xc, tc, nc = x.cuda(), t.cuda(), noise.cuda()
    unet._flat_grad.zero_()
    loss_full = gd.p_losses(xc, tc, nc)
    loss_full.backward()
    g_full = unet._flat_grad.clone()
    assert torch.isfinite(g_full).all() and g_full.abs().max().item() > 0
    unet._flat_grad.zero_()
    losses = []
    q = B // 4
    for i in range(4):
        s = slice(i * q, (i + 1) * q)
        loss = gd.p_losses(xc[s].contiguous(), tc[s].contiguous(), nc[s].contiguous())
        loss.backward()                                                      # gradients accumulate like torch autograd
        losses.append(loss.item())
    g_quarters = unet._flat_grad.clone() / 4
    assert abs(sum(losses) / 4 - loss_full.item()) <= 1e-4 * abs(loss_full.item())
    assert_close(g_quarters, g_full, "gradient of the batch mean vs mean of the quarter gradients", 1e-4)
    # anchor to the oracle: the loss of the three subset samples, same engine, against the CPU restatement
    with torch.no_grad():
        buf = O.diffusion_buffers(T)
        sub = O.p_losses(params, spec, buf, x[SUBSET], t[SUBSET], noise[SUBSET], "l2").item()
        got = gd.p_losses(xc[SUBSET].contiguous(), tc[SUBSET].contiguous(), nc[SUBSET].contiguous()).item()
    assert abs(got - sub) <= REL_TOL * abs(sub)
