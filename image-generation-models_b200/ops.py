"""torch.autograd wrappers over the generic operator entry points of libigm_b200 (csrc/ops.cu).

Tensors are logical NCHW kept in ``torch.channels_last`` memory format, i.e. physically the NHWC
layout the kernels use, so no layout copies happen between consecutive ops.  fp32 only, CUDA only.
"""
import ctypes as C

import torch

from . import _lib

CL = torch.channels_last


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _cl(x):
    if x.device.type != "cuda":
        raise RuntimeError("libigm_b200 runs on CUDA (B200, sm_100a) only; there is no CPU fallback")
    if x.dtype != torch.float32:
        raise TypeError("libigm_b200 computes in fp32")
    return x.contiguous(memory_format=CL)


def _ws(weight, geo):
    """Scratch of one conv call: packed weights, plus the bf16 operand staging of the tensor-core route when the geometry
    (B, H, W, Cin, Cout, KH, KW, stride, pad_h, pad_w, dil, transposed, OH, OW) qualifies for it."""
    lib = _lib.load()
    B, H, W, Cin, Cout, KH, KW, stride, pad_h, pad_w, dil, _transposed, OH, OW = geo
    n = lib.igm_conv2d_workspace_floats(B, H, W, Cin, Cout, KH, KW, stride, pad_h, pad_w, dil, OH, OW)
    return torch.empty(int(n), dtype=torch.float32, device=weight.device)


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual, stride, pad_h, pad_w, dil, transposed):
        lib = _lib.load()
        x = _cl(x)
        w = weight.contiguous()
        B, Cin, H, W = x.shape
        if transposed:
            Cout, KH, KW = w.shape[1], w.shape[2], w.shape[3]
            OH = (H - 1) * stride - 2 * pad_h + dil * (KH - 1) + 1
            OW = (W - 1) * stride - 2 * pad_w + dil * (KW - 1) + 1
            if dil != 1:
                raise NotImplementedError("dilated ConvTranspose2d")
        else:
            Cout, KH, KW = w.shape[0], w.shape[2], w.shape[3]
            OH = (H + 2 * pad_h - dil * (KH - 1) - 1) // stride + 1
            OW = (W + 2 * pad_w - dil * (KW - 1) - 1) // stride + 1
        y = torch.empty((B, Cout, OH, OW), device=x.device, dtype=torch.float32, memory_format=CL)
        geo = (B, H, W, Cin, Cout, KH, KW, stride, pad_h, pad_w, dil, int(transposed), OH, OW)
        b = None if bias is None else bias.contiguous()
        r = None
        if residual is not None:
            r = _cl(residual)
            if r.shape != y.shape:
                raise ValueError(f"residual shape {tuple(r.shape)} != conv output {tuple(y.shape)}")
        rc = lib.igm_conv2d_forward(_ptr(x), _ptr(w), _ptr(b), _ptr(r), _ptr(y), *geo, _ptr(_ws(w, geo)), _stream())
        _lib.check(None, rc)
        ctx.save_for_backward(x, w)
        ctx.geo = geo
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, w = ctx.saved_tensors
        dy = _cl(dy)
        dx = torch.empty_like(x, memory_format=CL) if ctx.needs_input_grad[0] else None
        dw = torch.zeros_like(w) if ctx.needs_input_grad[1] else None
        db = torch.zeros(ctx.geo[4], device=x.device) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        rc = lib.igm_conv2d_backward(_ptr(x), _ptr(w), _ptr(dy), _ptr(dx), _ptr(dw), _ptr(db), *ctx.geo, _ptr(_ws(w, ctx.geo)),
                                     _stream())
        _lib.check(None, rc)
        dres = dy if (ctx.has_res and ctx.needs_input_grad[3]) else None
        return dx, dw, db, dres, None, None, None, None, None


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _one(v):
    if isinstance(v, int):
        return v
    if v[0] != v[1]:
        raise NotImplementedError("anisotropic stride / dilation")
    return int(v[0])


def conv2d(x, weight, bias=None, stride=1, padding=0, dilation=1, residual=None):
    """F.conv2d (+ residual fused into the epilogue)."""
    ph, pw = _pair(padding)
    return _ConvFn.apply(x, weight, bias, residual, _one(stride), int(ph), int(pw), _one(dilation), False)


def conv_transpose2d(x, weight, bias=None, stride=1, padding=0, residual=None):
    ph, pw = _pair(padding)
    return _ConvFn.apply(x, weight, bias, residual, _one(stride), int(ph), int(pw), 1, True)


class _ActFn(torch.autograd.Function):
    """kind 0 relu, 1 elu, 2 tanh*sigmoid gate, 3 tanh*tanh gate (gates halve the channel count)."""

    @staticmethod
    def forward(ctx, x, kind, cond):
        lib = _lib.load()
        x = _cl(x)
        B, Cx, H, W = x.shape
        Cy = Cx // 2 if kind >= 2 else Cx
        if cond is not None:
            if kind < 2:
                raise ValueError("conditioning applies to the gated activations only")
            cond = cond.reshape(B, Cx).contiguous().float()
        y = torch.empty((B, Cy, H, W), device=x.device, dtype=torch.float32, memory_format=CL)
        rc = lib.igm_act_forward(kind, _ptr(x), _ptr(cond), H * W, _ptr(y), B * H * W, Cy, _stream())
        _lib.check(None, rc)
        ctx.kind = kind
        ctx.has_cond = cond is not None
        ctx.save_for_backward(y if kind == 0 else x, cond)
        ctx.shape = (B, Cx, Cy, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        ref, cond = ctx.saved_tensors
        B, Cx, Cy, H, W = ctx.shape
        dy = _cl(dy)
        dx = torch.empty((B, Cx, H, W), device=dy.device, dtype=torch.float32, memory_format=CL)
        dcond = torch.empty((B, Cx), device=dy.device) if (ctx.has_cond and ctx.needs_input_grad[2]) else None
        rc = lib.igm_act_backward(ctx.kind, _ptr(ref), _ptr(cond), H * W, _ptr(dy), _ptr(dx), _ptr(dcond), B * H * W, Cy,
                                  _stream())
        _lib.check(None, rc)
        return dx, None, dcond


def relu(x):
    return _ActFn.apply(x, 0, None)


def elu(x):
    return _ActFn.apply(x, 1, None)


def gate_tanh_sigmoid(x, cond=None):
    """tanh(x[:, :C] + cond[:, :C]) * sigmoid(x[:, C:] + cond[:, C:]); cond is [B, 2C] (per image) or None."""
    return _ActFn.apply(x, 2, cond)


def gate_tanh_tanh(x, cond=None):
    return _ActFn.apply(x, 3, cond)


class _STFn(torch.autograd.Function):
    """e + (q - e).detach(): the value keeps torch's two roundings, the gradient goes to e only."""

    @staticmethod
    def forward(ctx, e, q):
        lib = _lib.load()
        e = _cl(e)
        q = _cl(q)
        y = torch.empty_like(e, memory_format=CL)
        _lib.check(None, lib.igm_ewise(1, _ptr(e), _ptr(q), _ptr(y), e.numel(), _stream()))
        return y

    @staticmethod
    def backward(ctx, dy):
        return dy, None


def straight_through(e, q):
    return _STFn.apply(e, q)


class _MSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        lib = _lib.load()
        a = _cl(a) if a.dim() == 4 else a.contiguous()
        b = b.contiguous(memory_format=CL) if a.dim() == 4 else b.contiguous()
        if a.shape != b.shape or a.stride() != b.stride():
            raise ValueError("mse_loss operands must share shape and layout")
        loss = torch.empty(1, device=a.device)
        _lib.check(None, lib.igm_mse(_ptr(a), _ptr(b), a.numel(), _ptr(loss), None, None, _stream()))
        ctx.save_for_backward(a, b)
        return loss[0]

    @staticmethod
    def backward(ctx, dl):
        lib = _lib.load()
        a, b = ctx.saved_tensors
        da = torch.empty_like(a)
        d = dl.reshape(1).contiguous().float()
        _lib.check(None, lib.igm_mse(_ptr(a), _ptr(b), a.numel(), None, _ptr(d), _ptr(da), _stream()))
        return da, None


def mse_loss(a, b):
    """F.mse_loss(a, b) with the gradient flowing to ``a``."""
    return _MSEFn.apply(a, b)


class _CE256Fn(torch.autograd.Function):
    """nll[N, C, H, W] of 256-way logits given as the conv_out tensor [N, 256*C, H, W] (channel = cls*C + ch)."""

    @staticmethod
    def forward(ctx, logits, target):
        lib = _lib.load()
        logits = _cl(logits)
        N, K, H, W = logits.shape
        Cc = K // 256
        tgt = target.permute(0, 2, 3, 1).contiguous().to(torch.int64)     # [N, H, W, C]
        nll = torch.empty((N, H, W, Cc), device=logits.device, dtype=torch.float32)
        rc = lib.igm_ce256(_ptr(logits), _ptr(tgt), _ptr(nll), None, None, N * H * W, Cc, _stream())
        _lib.check(None, rc)
        ctx.save_for_backward(logits, tgt)
        return nll.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, d_nll):
        lib = _lib.load()
        logits, tgt = ctx.saved_tensors
        N, K, H, W = logits.shape
        Cc = K // 256
        d = d_nll.permute(0, 2, 3, 1).contiguous().float()
        dl = torch.empty_like(logits, memory_format=CL)
        rc = lib.igm_ce256(_ptr(logits), _ptr(tgt), None, _ptr(dl), _ptr(d), N * H * W, Cc, _stream())
        _lib.check(None, rc)
        return dl, None


def cross_entropy_256(logits, target):
    return _CE256Fn.apply(logits, target)
