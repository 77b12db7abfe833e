"""CPU restatement of the reference DDPM hot path (U-Net, diffusion, sampler).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker the CUDA path is
compared with, and the thing ``bench.py`` times as the CPU baseline.  It is a
plain fp32 (or fp64) PyTorch-functional restatement that takes the reference's
own ``state_dict`` (same keys, same NCHW shapes) and follows, line by line:

    /root/reference/src/models/ddpm.py
        SinusoidalPosEmb :47-59      Mish :62-64           Upsample :67-73
        Downsample :76-82            LayerNorm :85-95      Block :112-120
        ResnetBlock :123-143         LinearAttention :146-166
        Unet.__init__ :170-236       Unet.forward :238-261
        cosine_beta_schedule :281-291  GaussianDiffusion.__init__ :295-350
        predict_start_from_noise :359-364   q_posterior :367-376
        p_mean_variance :378-388     p_sample :390-397     p_sample_loop :399-409
        q_sample :433-444            p_losses :446-460

PARITY PINNING: the reference holds no golden vectors, so
this restatement is pinned against the reference itself, executed unmodified on
CPU through oracle/ref_loader.py: tests/test_oracle_vs_reference.py (runs where
/root/reference exists) and the committed fixtures under tests/golden/ made by
tests/golden/make_golden.py from the reference's own outputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

HEADS = 4          # reference ddpm.py:147
DIM_HEAD = 32      # reference ddpm.py:147
GROUPS = 8         # reference ddpm.py:113 (ResnetBlock ignores its own groups arg, :132-133)
GN_EPS = 1e-5      # torch.nn.GroupNorm default
LN_EPS = 1e-5      # reference ddpm.py:86


@dataclass
class UnetSpec:
    """Shape description of reference ``Unet(dim, channels, dim_mults)`` (ddpm.py:170-236)."""
    dim: int = 64
    channels: int = 3
    dim_mults: Tuple[int, ...] = (1, 2, 4, 8)

    @property
    def dims(self) -> List[int]:
        return [self.channels] + [self.dim * m for m in self.dim_mults]

    @property
    def in_out(self) -> List[Tuple[int, int]]:
        d = self.dims
        return list(zip(d[:-1], d[1:]))


def param_shapes(spec: UnetSpec) -> "Dict[str, Tuple[int, ...]]":
    """Ordered name -> shape map identical to reference ``Unet(...).state_dict()``.

    Order is registration order in ddpm.py:186-236 (ResnetBlock registers mlp,
    block1, block2, res_conv :125-134; Residual(PreNorm(LinearAttention))
    registers fn.fn.to_qkv, fn.fn.to_out, fn.norm :101-102, :151-152).
    """
    out: Dict[str, Tuple[int, ...]] = {}
    dim = spec.dim

    def lin(name, i, o):
        out[f"{name}.weight"] = (o, i)
        out[f"{name}.bias"] = (o,)

    def conv(name, i, o, k, bias=True):
        out[f"{name}.weight"] = (o, i, k, k)
        if bias:
            out[f"{name}.bias"] = (o,)

    def block(name, i, o):
        conv(f"{name}.block.0", i, o, 3)
        out[f"{name}.block.1.weight"] = (o,)
        out[f"{name}.block.1.bias"] = (o,)

    def resnet(name, i, o):
        lin(f"{name}.mlp.1", dim, o)
        block(f"{name}.block1", i, o)
        block(f"{name}.block2", o, o)
        if i != o:
            conv(f"{name}.res_conv", i, o, 1)

    def attn(name, c):
        conv(f"{name}.fn.fn.to_qkv", c, HEADS * DIM_HEAD * 3, 1, bias=False)
        conv(f"{name}.fn.fn.to_out", HEADS * DIM_HEAD, c, 1)
        out[f"{name}.fn.norm.g"] = (1, c, 1, 1)
        out[f"{name}.fn.norm.b"] = (1, c, 1, 1)

    lin("time_mlp.1", dim, dim * 4)
    lin("time_mlp.3", dim * 4, dim)
    in_out = spec.in_out
    n_res = len(in_out)
    for ind, (ci, co) in enumerate(in_out):
        resnet(f"downs.{ind}.0", ci, co)
        resnet(f"downs.{ind}.1", co, co)
        attn(f"downs.{ind}.2", co)
        if ind < n_res - 1:
            conv(f"downs.{ind}.3.conv", co, co, 3)
    # self.ups is registered (empty) right after self.downs (ddpm.py:195-196), so its
    # entries precede mid_block1/mid_attn/mid_block2 in state_dict order.
    for ind, (ci, co) in enumerate(reversed(in_out[1:])):
        resnet(f"ups.{ind}.0", co * 2, ci)
        resnet(f"ups.{ind}.1", ci, ci)
        attn(f"ups.{ind}.2", ci)
        # is_last is never true in the ups loop (ddpm.py:221-222): always Upsample.
        out[f"ups.{ind}.3.conv.weight"] = (ci, ci, 4, 4)   # ConvTranspose2d: (in, out, kh, kw)
        out[f"ups.{ind}.3.conv.bias"] = (ci,)
    mid = spec.dims[-1]
    resnet("mid_block1", mid, mid)
    attn("mid_attn", mid)
    resnet("mid_block2", mid, mid)
    block("final_conv.0", spec.dims[1], spec.dims[1])
    conv("final_conv.1", spec.dims[1], spec.channels, 1)
    return out


def init_params(spec: UnetSpec, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded synthetic parameters (NOT the reference's init; for fixtures/benches).

    Kaiming-uniform-like bounds so activations stay O(1) like the reference's
    default init; GroupNorm/LayerNorm gains perturbed away from 1/0 so that
    parity tests exercise the affine terms.
    """
    g = torch.Generator().manual_seed(seed)
    params = {}
    for name, shape in param_shapes(spec).items():
        if name.endswith("norm.g") or name.endswith("block.1.weight"):
            p = 1.0 + 0.1 * (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1)
        elif name.endswith("norm.b") or name.endswith("block.1.bias"):
            p = 0.1 * (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1)
        else:
            if len(shape) == 4:
                if ".3.conv.weight" in name and name.startswith("ups"):
                    fan_in = shape[1] * shape[2] * shape[3]   # torch: ConvTranspose fan_in uses dim 1
                else:
                    fan_in = shape[1] * shape[2] * shape[3]
            elif len(shape) == 2:
                fan_in = shape[1]
            else:
                # bias: bound from the matching weight's fan-in
                w = params[name[: -len("bias")] + "weight"]
                fan_in = w[0].numel()
            bound = 1.0 / math.sqrt(fan_in)
            p = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound
        params[name] = p.to(dtype)
    return params


# ----------------------------------------------------------------------------
# modules
# ----------------------------------------------------------------------------
def mish(x):
    """ddpm.py:62-64 — x * tanh(softplus(x)) (F.softplus beta=1, threshold=20)."""
    return x * torch.tanh(F.softplus(x))


def sinusoidal_pos_emb(time: torch.Tensor, dim: int) -> torch.Tensor:
    """ddpm.py:52-59.  Always computed in fp32 like the reference (arange/exp default dtype)."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, device=time.device) * -e)
    e = time[:, None] * e[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def layer_norm(x, g, b, eps=LN_EPS):
    """ddpm.py:92-95 — eps added to the std, not the variance."""
    std = torch.var(x, dim=1, unbiased=False, keepdim=True).sqrt()
    mean = torch.mean(x, dim=1, keepdim=True)
    return (x - mean) / (std + eps) * g + b


class _Taps:
    """Optional recorder of named intermediates (NCHW), for per-op GPU parity."""

    def __init__(self, enabled: bool):
        self.enabled = enabled
        self.t: Dict[str, torch.Tensor] = {}

    def __call__(self, name: str, v: torch.Tensor):
        if self.enabled:
            self.t[name] = v.detach().clone()
        return v


def _block(p, name, x, taps):
    """ddpm.py:115-120 — Conv3x3(pad 1) -> GroupNorm(8) -> Mish."""
    y = F.conv2d(x, p[f"{name}.block.0.weight"], p[f"{name}.block.0.bias"], padding=1)
    taps(f"{name}.conv", y)
    y = F.group_norm(y, GROUPS, p[f"{name}.block.1.weight"], p[f"{name}.block.1.bias"], GN_EPS)
    return mish(y)


def _resnet_block(p, name, x, t_emb, taps):
    """ddpm.py:136-143."""
    h = _block(p, f"{name}.block1", x, taps)
    h = h + F.linear(mish(t_emb), p[f"{name}.mlp.1.weight"], p[f"{name}.mlp.1.bias"])[:, :, None, None]
    taps(f"{name}.h1", h)
    h = _block(p, f"{name}.block2", h, taps)
    if f"{name}.res_conv.weight" in p:
        r = F.conv2d(x, p[f"{name}.res_conv.weight"], p[f"{name}.res_conv.bias"])
    else:
        r = x
    return taps(f"{name}.out", h + r)


def linear_attention(x, w_qkv, w_out, b_out):
    """ddpm.py:154-166 — softmax over the spatial axis of k, no q scaling."""
    b, c, h, w = x.shape
    qkv = F.conv2d(x, w_qkv)
    # 'b (qkv heads c) h w -> qkv b heads c (h w)'
    qkv = qkv.reshape(b, 3, HEADS, DIM_HEAD, h * w)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    k = k.softmax(dim=-1)
    context = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", context, q)
    out = out.reshape(b, HEADS * DIM_HEAD, h, w)
    return F.conv2d(out, w_out, b_out)


def _attn(p, name, x, taps):
    """Residual(PreNorm(dim, LinearAttention(dim))) — ddpm.py:39-45, :98-106."""
    n = layer_norm(x, p[f"{name}.fn.norm.g"], p[f"{name}.fn.norm.b"])
    taps(f"{name}.ln", n)
    a = linear_attention(n, p[f"{name}.fn.fn.to_qkv.weight"], p[f"{name}.fn.fn.to_out.weight"],
                         p[f"{name}.fn.fn.to_out.bias"])
    return taps(f"{name}.out", a + x)


def time_mlp(p, time: torch.Tensor, dim: int) -> torch.Tensor:
    """ddpm.py:188-193 — SinusoidalPosEmb -> Linear -> Mish -> Linear."""
    wdtype = p["time_mlp.1.weight"].dtype
    e = sinusoidal_pos_emb(time, dim).to(wdtype)
    e = F.linear(e, p["time_mlp.1.weight"], p["time_mlp.1.bias"])
    e = mish(e)
    return F.linear(e, p["time_mlp.3.weight"], p["time_mlp.3.bias"])


def unet_forward(p: Dict[str, torch.Tensor], spec: UnetSpec, x: torch.Tensor, time: torch.Tensor,
                 taps: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
    """ddpm.py:238-261.  ``p`` uses the reference Unet state_dict keys."""
    rec = _Taps(taps is not None)
    t = time_mlp(p, time, spec.dim)
    rec("time_mlp", t[:, :, None, None])
    h = []
    n_res = len(spec.in_out)
    for ind in range(n_res):
        x = _resnet_block(p, f"downs.{ind}.0", x, t, rec)
        x = _resnet_block(p, f"downs.{ind}.1", x, t, rec)
        x = _attn(p, f"downs.{ind}.2", x, rec)
        h.append(x)
        if ind < n_res - 1:
            x = F.conv2d(x, p[f"downs.{ind}.3.conv.weight"], p[f"downs.{ind}.3.conv.bias"], stride=2, padding=1)
            rec(f"downs.{ind}.3.out", x)
    x = _resnet_block(p, "mid_block1", x, t, rec)
    x = _attn(p, "mid_attn", x, rec)
    x = _resnet_block(p, "mid_block2", x, t, rec)
    for ind in range(n_res - 1):
        x = torch.cat((x, h.pop()), dim=1)
        x = _resnet_block(p, f"ups.{ind}.0", x, t, rec)
        x = _resnet_block(p, f"ups.{ind}.1", x, t, rec)
        x = _attn(p, f"ups.{ind}.2", x, rec)
        x = F.conv_transpose2d(x, p[f"ups.{ind}.3.conv.weight"], p[f"ups.{ind}.3.conv.bias"], stride=2, padding=1)
        rec(f"ups.{ind}.3.out", x)
    x = _block(p, "final_conv.0", x, rec)
    rec("final_conv.0.out", x)
    x = F.conv2d(x, p["final_conv.1.weight"], p["final_conv.1.bias"])
    if taps is not None:
        taps.update(rec.t)
    return x


# ----------------------------------------------------------------------------
# diffusion
# ----------------------------------------------------------------------------
def cosine_beta_schedule(timesteps: int, s: float = 0.008) -> np.ndarray:
    """ddpm.py:281-291 (float64 numpy; note linspace(0, steps, steps))."""
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    ac = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return np.clip(betas, a_min=0, a_max=0.999)


SCHEDULE_KEYS = (
    "betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
    "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
    "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
    "posterior_mean_coef1", "posterior_mean_coef2",
)


def linear_beta_schedule(timesteps: int) -> np.ndarray:
    """ddpm.py:275-279 — torch.linspace(scale * 1e-4, scale * 0.02, T, dtype=float64), scale = 1000 / T."""
    scale = 1000 / timesteps
    return torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64).numpy()


def diffusion_buffers(timesteps: int = 1000, betas=None) -> Dict[str, torch.Tensor]:
    """The 12 fp32 schedule buffers of GaussianDiffusion.__init__ (ddpm.py:317-350); ``betas`` overrides the
    cosine default like the constructor's ``betas=`` argument (:303-310)."""
    betas = cosine_beta_schedule(timesteps) if betas is None else np.asarray(betas, dtype=np.float64)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    pv = betas * (1.0 - ac_prev) / (1.0 - ac)
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    return {
        "betas": f32(betas),
        "alphas_cumprod": f32(ac),
        "alphas_cumprod_prev": f32(ac_prev),
        "sqrt_alphas_cumprod": f32(np.sqrt(ac)),
        "sqrt_one_minus_alphas_cumprod": f32(np.sqrt(1.0 - ac)),
        "log_one_minus_alphas_cumprod": f32(np.log(1.0 - ac)),
        "sqrt_recip_alphas_cumprod": f32(np.sqrt(1.0 / ac)),
        "sqrt_recipm1_alphas_cumprod": f32(np.sqrt(1.0 / ac - 1)),
        "posterior_variance": f32(pv),
        "posterior_log_variance_clipped": f32(np.log(np.maximum(pv, 1e-20))),
        "posterior_mean_coef1": f32(betas * np.sqrt(ac_prev) / (1.0 - ac)),
        "posterior_mean_coef2": f32((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)),
    }


def _extract(a, t, x):
    """ddpm.py:263-266."""
    return a.to(x.dtype).gather(-1, t).reshape(t.shape[0], *((1,) * (x.dim() - 1)))


def q_sample(buf, x_start, t, noise):
    """ddpm.py:433-444."""
    return (_extract(buf["sqrt_alphas_cumprod"], t, x_start) * x_start
            + _extract(buf["sqrt_one_minus_alphas_cumprod"], t, x_start) * noise)


def p_losses(p, spec, buf, x_start, t, noise, loss_type="l1"):
    """ddpm.py:446-460."""
    x_noisy = q_sample(buf, x_start, t, noise)
    x_recon = unet_forward(p, spec, x_noisy, t)
    if loss_type == "l1":
        return (noise - x_recon).abs().mean()
    if loss_type == "l2":
        return F.mse_loss(noise, x_recon)
    raise NotImplementedError(loss_type)


def p_sample(p, spec, buf, x, t, noise, clip_denoised=True):
    """ddpm.py:378-397 with the noise draw injected instead of torch.randn."""
    eps = unet_forward(p, spec, x, t)
    x_recon = (_extract(buf["sqrt_recip_alphas_cumprod"], t, x) * x
               - _extract(buf["sqrt_recipm1_alphas_cumprod"], t, x) * eps)
    if clip_denoised:
        x_recon = x_recon.clamp(-1.0, 1.0)
    mean = (_extract(buf["posterior_mean_coef1"], t, x) * x_recon
            + _extract(buf["posterior_mean_coef2"], t, x) * x)
    log_var = _extract(buf["posterior_log_variance_clipped"], t, x)
    nonzero_mask = (1 - (t == 0).to(x.dtype)).reshape(x.shape[0], *((1,) * (x.dim() - 1)))
    return mean + nonzero_mask * (0.5 * log_var).exp() * noise


@torch.no_grad()
def p_sample_loop(p, spec, buf, img, noises, t_start=None, n_steps=None):
    """ddpm.py:399-409 — ``img`` is x_T, ``noises[k]`` the draw used at the k-th executed step.

    Runs t = t_start, t_start-1, ... for n_steps steps (defaults: the full chain).
    """
    T = buf["betas"].shape[0]
    t_start = T - 1 if t_start is None else t_start
    n_steps = t_start + 1 if n_steps is None else n_steps
    b = img.shape[0]
    for k in range(n_steps):
        i = t_start - k
        img = p_sample(p, spec, buf, img, torch.full((b,), i, dtype=torch.long), noises[k])
    return img


def adam_step(params, grads, m, v, step, lr, b1, b2, eps=1e-8):
    """torch.optim.Adam (no weight decay / amsgrad) as called at ddpm.py:507-511."""
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    for k in params:
        g = grads[k]
        m[k].mul_(b1).add_(g, alpha=1 - b1)
        v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v[k].sqrt() / math.sqrt(bc2)).add_(eps)
        params[k].addcdiv_(m[k], denom, value=-lr / bc1)
