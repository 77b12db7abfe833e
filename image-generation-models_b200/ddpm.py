"""Host-side mirror of the reference DDPM classes, backed by libigm_b200.so.

Same class names, constructor signatures, method names and ``state_dict`` keys as
``/root/reference/src/models/ddpm.py`` (Unet :169-261, GaussianDiffusion :294-466,
DDPM :469-521) so that Hydra's ``_target_`` instantiation, Lightning's
``training_step`` / ``validation_step`` / ``configure_optimizers`` calls and
reference checkpoints keep working, while every FLOP of the U-Net, the loss, its
gradients, the Adam update and the T-step sampler runs in the hand-written sm_100a
kernels behind the C ABI (include/igm_b200.h).  PyTorch is used for device memory,
streams, RNG draws and (optionally) torch.distributed only.

There is NO eager / CPU fallback: the parameter-holder modules below have no
``forward`` of their own.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import nn

import os

from . import _lib

_EAGER_BWD = os.environ.get("IGM_EAGER_BWD", "1") != "0"
# Bucketed gradient all-reduce overlapped with backward: IGM_DDP_OVERLAP=1 / 0 forces it on / off; by default it is used
# from 64 MB of gradients on.  Measured on B200 (profiles/r2_scaling.md): NCCL's CTAs cannot co-reside with the persistent
# one-CTA-per-SM conv kernels, so the overlap is worth +0.4 ... +0.9 % on the 119 MB CelebA-64 arena and -0.3 ... 0 % on
# the 30.5 MB CIFAR-10 arena, where one all-reduce after the backward pass is as good.
_DDP_OVERLAP_ENV = os.environ.get("IGM_DDP_OVERLAP", "auto")


def _ddp_overlap(unet) -> bool:
    v = getattr(unet, "ddp_overlap", None)
    if v is not None:
        return bool(v)
    if _DDP_OVERLAP_ENV in ("0", "1"):
        return _DDP_OVERLAP_ENV == "1"
    return unet._flat.numel() * 4 >= (64 << 20)

try:  # the real Lightning base class when it is installed, a minimal stand-in otherwise
    from pytorch_lightning import LightningModule as _LightningModule  # type: ignore
    _HAVE_LIGHTNING = True
except Exception:  # pragma: no cover - Lightning is absent from this image
    _HAVE_LIGHTNING = False

    class _AttrDict(dict):
        __getattr__ = dict.__getitem__

        def __setattr__(self, k, v):
            self[k] = v

    class _LightningModule(nn.Module):
        """The slice of LightningModule the reference models use (hparams, log, device)."""

        def __init__(self):
            super().__init__()
            self.hparams = _AttrDict()
            self.logged: Dict[str, float] = {}

        def save_hyperparameters(self, **kw):
            self.hparams.update(kw)

        def log(self, name, value, *a, **k):
            self.logged[name] = value

        @property
        def device(self):
            p = next(self.parameters(), None)
            return p.device if p is not None else torch.device("cpu")


HEADS, DIM_HEAD = 4, 32   # reference ddpm.py:147


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"libigm_b200 computes in fp32; got {t.dtype}")
    return t.contiguous()


# ---------------------------------------------------------------------------
# parameter holders (state_dict-compatible module tree, no arithmetic)
# ---------------------------------------------------------------------------
class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: the arithmetic lives in libigm_b200 (no eager fallback)")


def _block(dim, dim_out):
    b = _Holder()
    b.block = nn.Sequential(nn.Conv2d(dim, dim_out, 3, padding=1), nn.GroupNorm(8, dim_out), _Holder())
    return b


def _resnet_block(dim, dim_out, time_emb_dim):
    r = _Holder()
    r.mlp = nn.Sequential(_Holder(), nn.Linear(time_emb_dim, dim_out))
    r.block1 = _block(dim, dim_out)
    r.block2 = _block(dim_out, dim_out)
    r.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()
    return r


def _attention(dim):
    la = _Holder()
    la.to_qkv = nn.Conv2d(dim, HEADS * DIM_HEAD * 3, 1, bias=False)
    la.to_out = nn.Conv2d(HEADS * DIM_HEAD, dim, 1)
    norm = _Holder()
    norm.g = nn.Parameter(torch.ones(1, dim, 1, 1))
    norm.b = nn.Parameter(torch.zeros(1, dim, 1, 1))
    pre = _Holder()
    pre.fn = la
    pre.norm = norm
    res = _Holder()
    res.fn = pre
    return res


def _resample(conv):
    h = _Holder()
    h.conv = conv
    return h


class _Engine:
    """One igm_ctx: plan + activations for (device, H, W, max_batch, training)."""

    def __init__(self, cfg: "_lib.UnetCfg", device: torch.device):
        self.lib = _lib.load()
        self.ctx = C.c_void_p()
        self.device = device
        idx = device.index if device.index is not None else torch.cuda.current_device()
        rc = self.lib.igm_unet_create(C.byref(self.ctx), C.byref(cfg), idx)
        if rc != 0:
            raise _lib.EngineError(f"igm_unet_create failed ({rc}): {self.lib.igm_last_error(None).decode()}")
        self.cfg = cfg
        self.sched_key = None
        self.packed_version = None
        self.bucket_ranges = None

    def check(self, rc):
        _lib.check(self.ctx, rc)

    def param_table(self):
        n = self.lib.igm_unet_num_params(self.ctx)
        out = []
        name = C.create_string_buffer(256)
        off, nd = C.c_int64(), C.c_int32()
        shape = (C.c_int64 * 4)()
        for i in range(n):
            self.check(self.lib.igm_unet_param_info(self.ctx, i, name, 256, C.byref(off), C.byref(nd), shape))
            out.append((name.value.decode(), off.value, tuple(shape[j] for j in range(nd.value))))
        return out, self.lib.igm_unet_param_elems(self.ctx)

    def close(self):
        if self.ctx:
            self.lib.igm_unet_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class Unet(nn.Module):
    """Drop-in for reference ``Unet`` (src/models/ddpm.py:169-261)."""

    def __init__(self, dim, out_dim=None, dim_mults=(1, 2, 4, 8), groups=8, channels=3, with_time_emb=True):
        super().__init__()
        if not with_time_emb:
            raise NotImplementedError("libigm_b200 implements the with_time_emb=True network the reference trains")
        if out_dim is not None and out_dim != channels:
            raise NotImplementedError("out_dim != channels is never used by the reference models")
        self.channels = channels
        self.dim = dim
        self.dim_mults = tuple(int(m) for m in dim_mults)
        dims = [channels, *[dim * m for m in self.dim_mults]]
        in_out = list(zip(dims[:-1], dims[1:]))
        time_dim = dim
        # construction order == the reference's, so torch's default init consumes the RNG identically
        self.time_mlp = nn.Sequential(_Holder(), nn.Linear(dim, dim * 4), _Holder(), nn.Linear(dim * 4, dim))
        self.downs = nn.ModuleList([])
        self.ups = nn.ModuleList([])   # registered before the mid blocks, like the reference (:195-196)
        n_res = len(in_out)
        for ind, (d_in, d_out) in enumerate(in_out):
            last = ind >= n_res - 1
            self.downs.append(nn.ModuleList([
                _resnet_block(d_in, d_out, time_dim),
                _resnet_block(d_out, d_out, time_dim),
                _attention(d_out),
                _resample(nn.Conv2d(d_out, d_out, 3, 2, 1)) if not last else nn.Identity(),
            ]))
        mid = dims[-1]
        self.mid_block1 = _resnet_block(mid, mid, time_dim)
        self.mid_attn = _attention(mid)
        self.mid_block2 = _resnet_block(mid, mid, time_dim)
        for ind, (d_in, d_out) in enumerate(reversed(in_out[1:])):
            # `is_last` is never true here in the reference (:221-222): every up stage upsamples
            self.ups.append(nn.ModuleList([
                _resnet_block(d_out * 2, d_in, time_dim),
                _resnet_block(d_in, d_in, time_dim),
                _attention(d_in),
                _resample(nn.ConvTranspose2d(d_in, d_in, 4, 2, 1)),
            ]))
        fin = _Holder()
        fin.block = nn.Sequential(nn.Conv2d(dims[1], dims[1], 3, padding=1), nn.GroupNorm(8, dims[1]), _Holder())
        self.final_conv = nn.Sequential(fin, nn.Conv2d(dims[1], channels, 1))

        self._engine: Optional[_Engine] = None
        self._flat: Optional[torch.Tensor] = None
        self._flat_grad: Optional[torch.Tensor] = None
        self._layout = None
        self._anchor: Optional[torch.Tensor] = None
        self._timesteps = 1000
        self._loss_type = 1
        self._flatten()

    # ---- flat parameter / gradient arenas ------------------------------------
    def _compute_layout(self):
        off, layout = 0, []
        for name, p in self.named_parameters():
            layout.append((name, off, tuple(p.shape)))
            off += (p.numel() + 3) // 4 * 4   # 16-byte alignment of every tensor (matches the C side)
        return layout, off

    @torch.no_grad()
    def _flatten(self):
        """Re-home every parameter (and its .grad) as a view into one flat fp32 arena."""
        layout, total = self._compute_layout()
        dev = next(self.parameters()).device
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        grad = torch.zeros(total, dtype=torch.float32, device=dev)
        offsets = {name: off for name, off, _ in layout}
        for mod_name, mod in self.named_modules():
            for key, p in list(mod._parameters.items()):
                if p is None:
                    continue
                off = offsets[f"{mod_name}.{key}" if mod_name else key]
                n = p.numel()
                flat[off:off + n].copy_(p.detach().reshape(-1))
                if p.grad is not None:
                    grad[off:off + n].copy_(p.grad.reshape(-1))
                # A Parameter built FROM the view shares the arena's version counter, so any
                # in-place edit (optimizer.step, load_state_dict) bumps flat._version -> repack.
                q = nn.Parameter(flat[off:off + n].view(p.shape), requires_grad=p.requires_grad)
                q.grad = grad[off:off + n].view(p.shape)
                mod._parameters[key] = q
        self._flat, self._flat_grad, self._layout = flat, grad, layout
        self._pend, self._pend_gen = None, 0
        self._zero_version = None
        self._fwd_gen = getattr(self, "_fwd_gen", 0) + 1
        self._synced = False
        self._comm = None
        self._anchor = torch.zeros(1, device=dev, requires_grad=True)
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        self._flatten()
        return self

    def attach_grads(self, zero: bool = False):
        """Make every ``p.grad`` a view of the flat gradient arena (after zero_grad(set_to_none=True))."""
        detached = False
        for (name, off, shape), p in zip(self._layout, self.parameters()):
            g = p.grad
            if g is None or g.data_ptr() != self._flat_grad.data_ptr() + 4 * off:
                if not detached and zero is False:
                    # a grad was dropped / replaced: start from a clean arena like torch would
                    self._flat_grad.zero_()
                detached = True
                p.grad = self._flat_grad[off:off + p.numel()].view(shape)
        if zero:
            self._flat_grad.zero_()
            # torch bumps the version counter on every in-place edit: while it stays at this value the arena is known to
            # be all zeros (the kernels' own writes do not bump it; _reduce_into_grad clears the mark after a backward)
            self._zero_version = self._flat_grad._version

    def _bind_grad_target(self, which: str):
        """Point the engine's weight-gradient kernels at ``.grad``'s arena ("grad") or at the pending arena of an eager
        backward ("pend").  A switch re-uploads the engine's pointer tables (synchronous, rare: a training loop stays
        in one mode)."""
        e = self._engine
        if e.grad_target == which:
            return
        if which == "pend" and (self._pend is None or self._pend.device != self._flat.device or
                                self._pend.numel() != self._flat.numel()):
            self._pend = torch.zeros_like(self._flat)
        arena = self._flat_grad if which == "grad" else self._pend
        e.check(e.lib.igm_unet_bind_params(e.ctx, _ptr(self._flat), _ptr(arena)))
        e.grad_target = which

    def sync_parameters(self, force: bool = False):
        """Data parallel: every replica starts from rank 0's parameters (what DistributedDataParallel does at wrap time;
        the mirror is never wrapped).  One broadcast of the flat arena, once per arena."""
        if _world() > 1 and getattr(self, "ddp_sync", True) and (force or not self._synced):
            import torch.distributed as dist
            dist.broadcast(self._flat, src=0)     # in place: bumps the arena's version counter -> weights are re-packed
            self._synced = True

    def mark_dirty(self):
        """Call after editing parameters through ``.data`` (which bypasses the version counter)."""
        if self._engine is not None:
            self._engine.packed_version = None

    # ---- engine ---------------------------------------------------------------
    def _get_engine(self, B: int, H: int, W: int, device: torch.device, training: bool) -> _Engine:
        if device.type != "cuda":
            raise RuntimeError("libigm_b200 runs on CUDA (B200, sm_100a) only; there is no CPU fallback")
        if self._flat.device != device:
            raise RuntimeError(f"parameters live on {self._flat.device}, input on {device}")
        e = self._engine
        if e is not None and (e.cfg.max_batch < B or e.cfg.height != H or e.cfg.width != W or
                              (training and not e.cfg.training) or e.cfg.timesteps != self._timesteps or
                              e.cfg.loss_type != self._loss_type):
            training = training or bool(e.cfg.training)
            B = max(B, e.cfg.max_batch)
            e.close()
            e = self._engine = None
        if e is None:
            cfg = _lib.UnetCfg()
            cfg.dim, cfg.channels, cfg.n_mults = self.dim, self.channels, len(self.dim_mults)
            for i, m in enumerate(self.dim_mults):
                cfg.dim_mults[i] = m
            cfg.height, cfg.width, cfg.max_batch = H, W, B
            cfg.timesteps, cfg.loss_type, cfg.training = self._timesteps, self._loss_type, int(training)
            e = _Engine(cfg, device)
            table, total = e.param_table()
            mine = [(n, o, s) for n, o, s in self._layout]
            if total != self._flat.numel() or [(n, o, tuple(s)) for n, o, s in table] != mine:
                e.close()
                raise RuntimeError("parameter layout of the module tree and of libigm_b200 disagree")
            e.check(e.lib.igm_unet_bind_params(e.ctx, _ptr(self._flat), _ptr(self._flat_grad)))
            e.grad_target = "grad"
            self._engine = e
            self.sync_parameters()
        if e.packed_version != self._flat._version:
            e.check(e.lib.igm_unet_pack_weights(e.ctx, _stream()))
            e.packed_version = self._flat._version
        return e

    def forward(self, x: torch.Tensor, time: torch.Tensor) -> torch.Tensor:
        """Unet.forward(x, time) — reference ddpm.py:238-261."""
        need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if need_grad:
            return _UnetFn.apply(self._anchor, self, x, time)
        return _unet_forward(self, x, time, training=False)

    def launch_count(self) -> int:
        return 0 if self._engine is None else int(self._engine.lib.igm_launch_count(self._engine.ctx))

    def profile_start(self):
        e = self._engine
        e.check(e.lib.igm_profile_start(e.ctx))

    def profile_stop(self):
        """-> {class: dict(launches, ms, flops, bytes)} for the launches since profile_start()."""
        e = self._engine
        buf = (_lib.ProfileEntry * 16)()
        n = e.lib.igm_profile_stop(e.ctx, buf, 16)
        if n < 0:
            e.check(n)
        return {buf[i].name.decode(): dict(launches=buf[i].launches, ms=buf[i].ms, flops=buf[i].flops,
                                           bytes=buf[i].bytes) for i in range(n)}

    def read_tap(self, name: str) -> torch.Tensor:
        """Named intermediate of the last forward as NCHW (tests / profiling)."""
        e = self._engine
        n = e.lib.igm_debug_read_tap(e.ctx, name.encode(), None, 0, _stream())
        if n < 0:
            e.check(int(n))
        out = torch.empty(int(n), dtype=torch.float32, device=self._flat.device)
        r = e.lib.igm_debug_read_tap(e.ctx, name.encode(), _ptr(out), n, _stream())
        if r < 0:
            e.check(int(r))
        return out


def _unet_forward(unet: Unet, x, time, training):
    x = _f32c(x)
    B, Cc, H, W = x.shape
    if Cc != unet.channels:
        raise ValueError(f"expected {unet.channels} channels, got {Cc}")
    t = time.to(torch.int64).contiguous()
    e = unet._get_engine(B, H, W, x.device, training)
    out = torch.empty_like(x)
    e.check(e.lib.igm_unet_forward(e.ctx, _ptr(x), _ptr(t), _ptr(out), B, _stream()))
    unet._fwd_gen += 1     # the engine keeps the activations of its LAST forward only (see _check_gen)
    return out


def _check_gen(unet: Unet, gen: int):
    """The engine stores one forward's activations.  A backward whose forward is no longer the engine's last one
    (two losses summed before one backward, a sampler / inference call in between) would silently differentiate the
    wrong activations: refuse instead."""
    if gen != unet._fwd_gen:
        raise RuntimeError("libigm_b200 keeps the activations of the most recent forward only: call backward() on a loss "
                           "before running another forward / sampler step through the same Unet")


def _reduce_into_grad(unet: Unet, e: "_Engine", run_backward, d_scale: float = 1.0, scalable: bool = False):
    """Backward into ``.grad`` with torch's accumulate semantics.  ``run_backward(scale)`` enqueues the kernels with the
    seed gradient multiplied by ``scale``.  One rank: the kernels add straight into the ``.grad`` arena.  Data parallel:
    this backward's gradients go to the pending arena, ONLY that delta is all-reduced (reducing the accumulated ``.grad``
    arena would re-reduce earlier micro-batches), and ``.grad += delta / world``.  When ``.grad`` is known to be all zeros
    (``zero_grad()`` just ran, the common loop) the delta IS the arena: the kernels write into it directly with the seed
    scaled by 1 / world and the arena is all-reduced in place -- no memset of the pending arena, no axpy pass."""
    sync = _world() > 1 and getattr(unet, "ddp_sync", True)
    overlap = _ddp_overlap(unet)
    if not sync:
        unet._bind_grad_target("grad")
        run_backward(1.0)
        unet._zero_version = None
        return
    zv = unet._zero_version
    if scalable and zv is not None and zv == unet._flat_grad._version and unet._flat_grad.is_cuda:
        unet._bind_grad_target("grad")
        run_backward(d_scale / _world())
        if overlap:
            _allreduce_overlapped(unet, e, unet._flat_grad)
        else:
            _allreduce(unet, unet._flat_grad)
        unet._zero_version = None
        return
    unet._bind_grad_target("pend")
    unet._pend.zero_()
    run_backward(1.0)
    if unet._pend.is_cuda and overlap:
        _allreduce_overlapped(unet, e, unet._pend)
    else:
        _allreduce(unet, unet._pend)
    e.check(e.lib.igm_grad_axpy(e.ctx, _ptr(unet._flat_grad), _ptr(unet._pend), None, C.c_float(d_scale / _world()),
                                unet._flat_grad.numel(), _stream()))
    unet._pend_gen += 1    # the pending arena no longer holds an eager training_step's gradients
    unet._zero_version = None


class _UnetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, unet, x, time):
        ctx.unet = unet
        ctx.need_dx = x.requires_grad
        out = _unet_forward(unet, x, time, training=True)
        ctx.gen = unet._fwd_gen
        return out

    @staticmethod
    def backward(ctx, d_out):
        unet = ctx.unet
        e = unet._engine
        _check_gen(unet, ctx.gen)
        unet.attach_grads()
        d_out = _f32c(d_out)
        dx = torch.empty_like(d_out) if ctx.need_dx else None
        # parameter gradients are the MEAN over ranks (like DistributedDataParallel); dx stays this rank's own
        _reduce_into_grad(unet, e, lambda scale: e.check(e.lib.igm_unet_backward(e.ctx, _ptr(d_out), _ptr(dx), _stream())))
        return None, None, dx, None


def _world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    return 1


def _allreduce_overlapped(unet: Unet, e: "_Engine", arena: torch.Tensor):
    """The exchange of the backward pass that was JUST enqueued, overlapped with it: the engine completes the gradient
    arena bucket by bucket (igm_unet_grad_buckets: ups/mid/final first, downs.(n-1) .. downs.1, time_mlp + downs.0 last);
    a communication stream waits for bucket k's events (igm_unet_bucket_wait, device-side) and all-reduces that range
    while the compute stream is still running the rest of the backward pass.  The compute stream then waits for the
    communication stream, so whatever follows (axpy into .grad, Adam) sees the reduced arena."""
    import torch.distributed as dist
    if e.bucket_ranges is None:
        lo, hi = (C.c_int64 * 16)(), (C.c_int64 * 16)()
        n = e.lib.igm_unet_grad_buckets(e.ctx, lo, hi, 16)
        if n < 1:
            e.check(n if n < 0 else -1)
        e.bucket_ranges = [(int(lo[k]), int(hi[k])) for k in range(n)]
    dev = arena.device
    if unet._comm is None or unet._comm.device != dev:
        unet._comm = torch.cuda.Stream(device=dev)
    comm, main = unet._comm, torch.cuda.current_stream(dev)
    with torch.cuda.stream(comm):
        for k, (lo_k, hi_k) in enumerate(e.bucket_ranges):
            e.check(e.lib.igm_unet_bucket_wait(e.ctx, k, C.c_void_p(comm.cuda_stream)))
            dist.all_reduce(arena[lo_k:hi_k], op=dist.ReduceOp.SUM)
    main.wait_stream(comm)


def _allreduce(unet: Unet, arena: torch.Tensor):
    """Data-parallel exchange of ONE backward's gradients (SUM over ranks) over the flat fp32 arena, in
    ``unet.ddp_buckets`` contiguous pieces (NCCL pipelines them; one piece = one ncclAllReduce)."""
    import torch.distributed as dist
    nb = max(1, int(getattr(unet, "ddp_buckets", 1)))
    if nb == 1:
        dist.all_reduce(arena, op=dist.ReduceOp.SUM)
        return
    n = arena.numel()
    step = (n + nb - 1) // nb
    step = (step + 1023) // 1024 * 1024
    for off in range(0, n, step):
        dist.all_reduce(arena[off:off + step], op=dist.ReduceOp.SUM)


# ---------------------------------------------------------------------------
# diffusion
# ---------------------------------------------------------------------------
def _cosine_betas(timesteps: int, s: float = 0.008) -> np.ndarray:
    """Cosine schedule with the reference's grid ``linspace(0, T+1, T+1)`` (ddpm.py:281-291), float64."""
    n = timesteps + 1
    grid = np.linspace(0, n, n)
    f = np.cos((grid / n + s) / (1 + s) * np.pi * 0.5) ** 2
    f = f / f[0]
    return np.clip(1 - f[1:] / f[:-1], a_min=0, a_max=0.999)


def cosine_beta_schedule(timesteps, s=0.008):
    """reference ddpm.py:281-291 (numpy float64)."""
    return _cosine_betas(timesteps, s)


def linear_beta_schedule(timesteps):
    """reference ddpm.py:275-279 (torch float64); pass it as ``GaussianDiffusion(betas=...)``."""
    scale = 1000 / timesteps
    return torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64)


def extract(a, t, x_shape):
    """reference ddpm.py:263-266."""
    b, *_ = t.shape
    return a.gather(-1, t).reshape(b, *((1,) * (len(x_shape) - 1)))


def noise_like(shape, device, repeat=False):
    """reference ddpm.py:268-273."""
    if repeat:
        return torch.randn((1, *shape[1:]), device=device).repeat(shape[0], *((1,) * (len(shape) - 1)))
    return torch.randn(shape, device=device)


_SCHED_FIELDS = ("sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
                 "sqrt_recipm1_alphas_cumprod", "posterior_log_variance_clipped", "posterior_mean_coef1",
                 "posterior_mean_coef2")


class GaussianDiffusion(nn.Module):
    """Drop-in for reference ``GaussianDiffusion`` (src/models/ddpm.py:294-466)."""

    def __init__(self, denoise_fn: Unet, *, image_size, channels=3, timesteps=1000, loss_type="l1", betas=None):
        super().__init__()
        if loss_type not in ("l1", "l2"):
            raise NotImplementedError(loss_type)
        self.channels = channels
        self.image_size = image_size
        self.denoise_fn = denoise_fn
        self.eager_backward = False   # True inside DDPM.training_step: see _PLossesFn
        self._rb, self._rb_pending = None, False
        if betas is not None:
            betas = betas.detach().cpu().numpy() if isinstance(betas, torch.Tensor) else np.asarray(betas)
        else:
            betas = _cosine_betas(timesteps)
        betas = betas.astype(np.float64)
        alphas = 1.0 - betas
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1.0, ac[:-1])
        self.num_timesteps = int(betas.shape[0])
        self.loss_type = loss_type
        post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
        # the reference's 12 buffers, same names / order / fp32 rounding (ddpm.py:325-346)
        for name, val in (
            ("betas", betas), ("alphas_cumprod", ac), ("alphas_cumprod_prev", ac_prev),
            ("sqrt_alphas_cumprod", np.sqrt(ac)), ("sqrt_one_minus_alphas_cumprod", np.sqrt(1.0 - ac)),
            ("log_one_minus_alphas_cumprod", np.log(1.0 - ac)), ("sqrt_recip_alphas_cumprod", np.sqrt(1.0 / ac)),
            ("sqrt_recipm1_alphas_cumprod", np.sqrt(1.0 / ac - 1)), ("posterior_variance", post_var),
            ("posterior_log_variance_clipped", np.log(np.maximum(post_var, 1e-20))),
            ("posterior_mean_coef1", betas * np.sqrt(ac_prev) / (1.0 - ac)),
            ("posterior_mean_coef2", (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)),
        ):
            self.register_buffer(name, torch.tensor(val, dtype=torch.float32))
        denoise_fn._timesteps = self.num_timesteps
        denoise_fn._loss_type = 1 if loss_type == "l1" else 2

    # ---- loss read-back that does not queue behind the eager backward ----------------
    def _start_loss_readback(self, loss: torch.Tensor):
        """Copy the scalar loss to pinned host memory on a side stream as soon as the forward has produced it; a
        ``loss.item()`` on the compute stream would wait for the backward kernels enqueued behind the forward."""
        dev = loss.device
        if self._rb is None or self._rb[0] != dev:
            self._rb = (dev, torch.cuda.Stream(device=dev), torch.empty(1, dtype=torch.float32).pin_memory(),
                        torch.cuda.Event(), torch.cuda.Event())
        _, side, host, ready, done = self._rb
        ready.record(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            side.wait_event(ready)
            host.copy_(loss.detach().reshape(1), non_blocking=True)
            done.record(side)
        loss.record_stream(side)
        self._rb_pending = True

    def loss_value(self, loss: torch.Tensor) -> float:
        """``loss.item()`` of the most recent eager p_losses (4 bytes device->host per step either way)."""
        if not self._rb_pending:
            return loss.item()
        self._rb_pending = False
        self._rb[4].synchronize()
        return float(self._rb[2][0])

    # ---- engine plumbing --------------------------------------------------------
    def _engine(self, B, H, W, device, training) -> _Engine:
        e = self.denoise_fn._get_engine(B, H, W, device, training)
        key = tuple(getattr(self, f).data_ptr() for f in _SCHED_FIELDS)
        if e.sched_key != key:
            s = _lib.Schedule()
            for f in _SCHED_FIELDS:
                buf = getattr(self, f)
                if buf.device != device or buf.dtype != torch.float32 or not buf.is_contiguous():
                    raise RuntimeError(f"schedule buffer {f} must be contiguous fp32 on {device}")
                setattr(s, f, buf.data_ptr())
            e.check(e.lib.igm_ddpm_set_schedule(e.ctx, C.byref(s)))
            e.sched_key = key
        return e

    @staticmethod
    def _extract(a, t, x):
        return a.gather(-1, t).reshape(t.shape[0], *((1,) * (x.dim() - 1)))

    # ---- thin public helpers (same names as the reference) ----------------------
    def q_mean_variance(self, x_start, t):
        mean = self._extract(self.sqrt_alphas_cumprod, t, x_start) * x_start
        variance = self._extract(1.0 - self.alphas_cumprod, t, x_start)
        log_variance = self._extract(self.log_one_minus_alphas_cumprod, t, x_start)
        return mean, variance, log_variance

    def predict_start_from_noise(self, x_t, t, noise):
        return (self._extract(self.sqrt_recip_alphas_cumprod, t, x_t) * x_t
                - self._extract(self.sqrt_recipm1_alphas_cumprod, t, x_t) * noise)

    def q_posterior(self, x_start, x_t, t):
        mean = (self._extract(self.posterior_mean_coef1, t, x_t) * x_start
                + self._extract(self.posterior_mean_coef2, t, x_t) * x_t)
        return (mean, self._extract(self.posterior_variance, t, x_t),
                self._extract(self.posterior_log_variance_clipped, t, x_t))

    def p_mean_variance(self, x, t, clip_denoised: bool):
        with torch.no_grad():
            eps = self.denoise_fn(x, t)
        x_recon = self.predict_start_from_noise(x, t=t, noise=eps)
        if clip_denoised:
            x_recon.clamp_(-1.0, 1.0)
        return self.q_posterior(x_start=x_recon, x_t=x, t=t)

    def q_sample(self, x_start, t, noise=None):
        """q(x_t | x_0) — reference ddpm.py:433-444 (one fused kernel)."""
        if noise is None:
            noise = torch.randn_like(x_start)
        x_start, noise = _f32c(x_start), _f32c(noise)
        B, _, H, W = x_start.shape
        e = self._engine(B, H, W, x_start.device, training=False)
        out = torch.empty_like(x_start)
        e.check(e.lib.igm_ddpm_q_sample(e.ctx, _ptr(x_start), _ptr(t.to(torch.int64).contiguous()), _ptr(noise),
                                        _ptr(out), B, _stream()))
        return out

    # ---- sampling ----------------------------------------------------------------
    @torch.no_grad()
    def _run_sampler(self, img, t_start, n_steps, noise=None, seed=0, clip_denoised=True):
        img = _f32c(img)
        B, _, H, W = img.shape
        e = self._engine(B, H, W, img.device, training=False)
        if noise is not None:
            noise = _f32c(noise)
            if noise.shape != (n_steps, *img.shape):
                raise ValueError("injected noise must be [n_steps, B, C, H, W]")
        e.check(e.lib.igm_ddpm_sample_loop(e.ctx, _ptr(img), _ptr(noise), C.c_uint64(seed), B, int(t_start),
                                           int(n_steps), int(bool(clip_denoised)), _stream()))
        self.denoise_fn._fwd_gen += 1   # the sampler's forwards overwrote the activations of any earlier forward
        return img

    @torch.no_grad()
    def p_sample(self, x, t, clip_denoised=True, repeat_noise=False):
        """One reverse step (reference ddpm.py:390-397); the draw comes from torch's generator
        exactly like the reference's ``noise_like`` (:268-273), then the fused kernel does the rest."""
        b = x.shape[0]
        if repeat_noise:
            noise = torch.randn((1, *x.shape[1:]), device=x.device).repeat(b, *((1,) * (x.dim() - 1)))
        else:
            noise = torch.randn(x.shape, device=x.device)
        ts = t.reshape(-1)
        t0 = int(ts[0].item())
        if not bool((ts == t0).all()):
            raise NotImplementedError("p_sample expects a batch-uniform timestep, as p_sample_loop provides")
        out = x.clone()
        return self._run_sampler(out, t0, 1, noise=noise[None], clip_denoised=clip_denoised)

    @torch.no_grad()
    def p_sample_loop(self, shape, noise=None, seed=None):
        """Full reverse chain (reference ddpm.py:399-409) as ONE C call: the per-step launch
        sequence is captured in a CUDA graph and replayed T times, no Python per step.
        ``noise`` ([T,B,C,H,W]) injects the draws (parity tests); otherwise the in-kernel
        Philox stream keyed by ``seed`` (drawn from torch's generator when None) is used."""
        device = self.betas.device
        img = torch.randn(shape, device=device)
        if seed is None and noise is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        return self._run_sampler(img, self.num_timesteps - 1, self.num_timesteps, noise=noise, seed=seed or 0)

    @torch.no_grad()
    def sample(self, batch_size=16):
        return self.p_sample_loop((batch_size, self.channels, *self.image_size))

    @torch.no_grad()
    def interpolate(self, x1, x2, t=None, weight=0.5):
        """reference ddpm.py:417-431."""
        b = x1.shape[0]
        t = self.num_timesteps - 1 if t is None else t
        assert x1.shape == x2.shape
        tb = torch.full((b,), t, device=x1.device, dtype=torch.long)
        xt1, xt2 = self.q_sample(x1, t=tb), self.q_sample(x2, t=tb)
        img = (1 - weight) * xt1 + weight * xt2
        if t == 0:
            return img
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        return self._run_sampler(img, t - 1, t, seed=seed)

    # ---- training -------------------------------------------------------------------
    def p_losses(self, x_start, t, noise=None):
        """reference ddpm.py:446-460: q_sample + U-Net + l1/l2 loss, one fused C call."""
        if noise is None:
            noise = torch.randn_like(x_start)
        if torch.is_grad_enabled():
            return _PLossesFn.apply(self.denoise_fn._anchor, self, x_start, t, noise, bool(self.eager_backward))
        return _p_losses_forward(self, x_start, t, noise, training=False)

    def forward(self, x, *args, **kwargs):
        b = x.shape[0]
        t = torch.randint(0, self.num_timesteps, (b,), device=x.device).long()
        return self.p_losses(x, t, *args, **kwargs)


def _p_losses_forward(gd: GaussianDiffusion, x_start, t, noise, training):
    x_start, noise = _f32c(x_start), _f32c(noise)
    B, _, H, W = x_start.shape
    e = gd._engine(B, H, W, x_start.device, training)
    loss = torch.empty((), dtype=torch.float32, device=x_start.device)
    e.check(e.lib.igm_ddpm_p_losses(e.ctx, _ptr(x_start), _ptr(t.to(torch.int64).contiguous()), _ptr(noise),
                                    _ptr(loss), B, _stream()))
    gd.denoise_fn._fwd_gen += 1
    return loss


class _PLossesFn(torch.autograd.Function):
    """Loss forward + backward.  ``eager`` (set by DDPM.training_step): the backward kernels (and the gradient
    all-reduce) are enqueued right behind the forward, with d_loss = 1, into the pending arena ``unet._pend`` — so the
    host's ``loss.item()`` read-back that the reference does between forward and backward (ddpm.py:499) no longer leaves
    the GPU idle while Python walks back into autograd.  ``backward`` is then ``.grad += d_loss * pending``: same
    accumulate semantics, and a ``zero_grad()`` between training_step and backward (Lightning's closure order) is safe."""

    @staticmethod
    def forward(ctx, anchor, gd, x_start, t, noise, eager):
        ctx.gd = gd
        ctx.eager = eager
        unet = gd.denoise_fn
        if not eager:
            loss = _p_losses_forward(gd, x_start, t, noise, training=True)
            ctx.fgen = unet._fwd_gen
            return loss
        x_start, noise = _f32c(x_start), _f32c(noise)
        B, _, H, W = x_start.shape
        e = gd._engine(B, H, W, x_start.device, True)
        unet._bind_grad_target("pend")
        loss = torch.empty((), dtype=torch.float32, device=x_start.device)
        e.check(e.lib.igm_ddpm_p_losses(e.ctx, _ptr(x_start), _ptr(t.to(torch.int64).contiguous()), _ptr(noise),
                                        _ptr(loss), B, _stream()))
        unet._fwd_gen += 1
        gd._start_loss_readback(loss)   # D2H of the scalar on a copy stream, ahead of the backward kernels
        unet._pend.zero_()
        sync = _world() > 1 and getattr(unet, "ddp_sync", True)
        e.check(e.lib.igm_ddpm_p_losses_backward(e.ctx, None, C.c_float(1.0 / _world() if sync else 1.0), _stream()))
        if sync:
            if _ddp_overlap(unet):
                _allreduce_overlapped(unet, e, unet._pend)
            else:
                _allreduce(unet, unet._pend)
        unet._pend_gen += 1
        ctx.gen = unet._pend_gen
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        unet = ctx.gd.denoise_fn
        e = unet._engine
        unet.attach_grads()
        d_loss = d_loss.to(torch.float32).contiguous()
        if ctx.eager:
            if ctx.gen != unet._pend_gen:
                raise RuntimeError("the eagerly computed gradients of this loss were overwritten by a later "
                                   "training_step; call backward() before the next one")
            e.check(e.lib.igm_grad_axpy(e.ctx, _ptr(unet._flat_grad), _ptr(unet._pend), _ptr(d_loss), C.c_float(1.0),
                                        unet._flat_grad.numel(), _stream()))
            unet._zero_version = None
            return None, None, None, None, None, None
        _check_gen(unet, ctx.fgen)
        _reduce_into_grad(unet, e, lambda scale: e.check(e.lib.igm_ddpm_p_losses_backward(e.ctx, _ptr(d_loss), C.c_float(scale),
                                                                                          _stream())), scalable=True)
        return None, None, None, None, None, None


# ---------------------------------------------------------------------------
# optimiser: torch.optim.Adam semantics, one fused kernel over the flat arenas
# ---------------------------------------------------------------------------
class FusedAdam(torch.optim.Optimizer):
    """Adam exactly as configured at reference ddpm.py:502-512, stepping ``unet``'s flat
    parameter arena with a single kernel (and re-packing the conv weights afterwards)."""

    def __init__(self, unet: Unet, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.unet = unet
        # the full torch.optim.Adam group (its other options at their defaults) so that state_dict()s interchange
        super().__init__(list(unet.parameters()), dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False,
                                                       maximize=False, foreach=None, capturable=False,
                                                       differentiable=False, fused=None, decoupled_weight_decay=False))
        self._step = 0
        self._m = None
        self._v = None

    def zero_grad(self, set_to_none: bool = False):
        # keep .grad views attached to the arena; zeroing is one memset
        self.unet.attach_grads(zero=True)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        u = self.unet
        if u._engine is None:
            raise RuntimeError("FusedAdam.step() before any forward/backward")
        if self._m is None or self._m.device != u._flat.device or self._m.numel() != u._flat.numel():
            self._m = torch.zeros_like(u._flat)
            self._v = torch.zeros_like(u._flat)
        g = self.param_groups[0]
        self._step += 1
        e = u._engine
        e.check(e.lib.igm_adam_step(e.ctx, _ptr(u._flat), _ptr(u._flat_grad), _ptr(self._m), _ptr(self._v),
                                    u._flat.numel(), g["lr"], g["betas"][0], g["betas"][1], g["eps"], self._step,
                                    1.0, _stream()))
        e.packed_version = None   # parameters changed behind torch's version counter
        return loss

    def state_dict(self):
        """torch.optim.Adam's format: per-parameter ``exp_avg`` / ``exp_avg_sq`` / ``step`` entries (views of the flat
        moment arenas), so a checkpoint written here resumes under the reference's ``torch.optim.Adam`` and vice versa."""
        sd = super().state_dict()
        if self._m is not None:
            step = torch.tensor(float(self._step))
            state = {}
            for i, ((name, off, shape), p) in enumerate(zip(self.unet._layout, self.unet.parameters())):
                n = p.numel()
                state[i] = {"step": step.clone(), "exp_avg": self._m[off:off + n].view(shape),
                            "exp_avg_sq": self._v[off:off + n].view(shape)}
            sd["state"] = state
        return sd

    def load_state_dict(self, sd):
        sd = dict(sd)
        legacy = sd.pop("igm", None)    # round-1 checkpoints kept the flat arenas under a private key
        state = sd.get("state") or {}
        super().load_state_dict({**sd, "state": {}})
        u = self.unet
        if legacy and legacy.get("exp_avg") is not None:
            self._step = int(legacy["step"])
            self._m = legacy["exp_avg"].to(u._flat.device, torch.float32).clone()
            self._v = legacy["exp_avg_sq"].to(u._flat.device, torch.float32).clone()
            return
        if not state:
            self._step, self._m, self._v = 0, None, None
            return
        self._m = torch.zeros_like(u._flat)
        self._v = torch.zeros_like(u._flat)
        steps = set()
        with torch.no_grad():
            for i, (name, off, shape) in enumerate(u._layout):
                st = state.get(i, state.get(str(i)))
                if st is None:
                    continue
                n = st["exp_avg"].numel()
                self._m[off:off + n].copy_(st["exp_avg"].reshape(-1))
                self._v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"FusedAdam steps every parameter together; the checkpoint holds different step counts {sorted(steps)}")
        self._step = steps.pop() if steps else 0


# ---------------------------------------------------------------------------
# LightningModule surface
# ---------------------------------------------------------------------------
@dataclass
class ValidationResult:
    """reference src/models/base.py:7-14."""
    others: dict = field(default_factory=dict)
    real_image: torch.Tensor = None
    fake_image: torch.Tensor = None
    recon_image: torch.Tensor = None
    label: torch.Tensor = None
    encode_latent: torch.Tensor = None


class DDPM(_LightningModule):
    """Drop-in for reference ``DDPM`` (src/models/ddpm.py:469-521)."""

    def __init__(self, datamodule, hidden_dim: int = 64, timesteps: int = 1000, loss_type: str = "l1",
                 dim_mults: Tuple[int] = (1, 2, 4, 8), lr: float = 0.0002, b1: float = 0.5, b2: float = 0.999,
                 optim="adam", **kwargs):
        super().__init__()
        # BaseModel.__init__ (reference src/models/base.py:17-27)
        self.width = datamodule.width
        self.height = datamodule.height
        self.channels = datamodule.channels
        self.input_normalize = datamodule.transforms.normalize
        self.output_act = "tanh" if self.input_normalize else "sigmoid"
        if _HAVE_LIGHTNING:
            self.save_hyperparameters(ignore=["datamodule"])
        else:
            self.save_hyperparameters(hidden_dim=hidden_dim, timesteps=timesteps, loss_type=loss_type,
                                      dim_mults=tuple(dim_mults), lr=lr, b1=b1, b2=b2, optim=optim, **kwargs)
        self.denoising_model = Unet(dim=hidden_dim, channels=self.channels, dim_mults=tuple(dim_mults))
        self.diffusion_model = GaussianDiffusion(self.denoising_model, image_size=(self.height, self.width),
                                                 timesteps=timesteps, loss_type=loss_type, channels=self.channels)

    def training_step(self, batch, batch_idx):
        imgs, _ = batch
        # forward AND backward kernels go out before the loss is read back (see _PLossesFn); IGM_EAGER_BWD=0 disables
        self.diffusion_model.eager_backward = _EAGER_BWD
        try:
            loss = self.diffusion_model(imgs)
        finally:
            self.diffusion_model.eager_backward = False
        self.log("train_loss/loss", self.diffusion_model.loss_value(loss))
        return loss

    def configure_optimizers(self):
        return FusedAdam(self.denoising_model, lr=self.hparams.lr, betas=(self.hparams.b1, self.hparams.b2))

    def validation_step(self, batch, batch_idx):
        imgs, labels = batch
        n = imgs.shape[0]
        t = torch.full((n,), self.hparams.timesteps - 1, device=imgs.device, dtype=torch.long)
        diffusion_imgs = self.diffusion_model.q_sample(imgs, t=t)
        fake = self.diffusion_model.sample(64) if batch_idx == 0 else None
        return ValidationResult(real_image=imgs, fake_image=fake, others={"diffusion": diffusion_imgs})
