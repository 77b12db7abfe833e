"""Linear attention (reference src/models/ddpm.py:154-166) with every per-pixel contraction written as a 1x1 convolution
whose weights differ per image - the operation csrc/conv_tc.cu already runs on the tensor cores for the sampler's M_b
shortcut (tc_plan_img).  Executable algebra for the planned training path; not a kernel and
not on any product path.  tests/test_attn_conv_model.py checks it against autograd of the reference formulation.

NHWC view, heads = 4, d = 32: Q, K, V, dOut are [B, n, 128] with channel (h, d); a per-image weight matrix is
Wt[b][c_out][c_in] (what the conv engine's weight tensor [B * N rows][K] holds), block-diagonal over heads.

  forward    P    = softmax_n(K)                               elementwise + the existing pixel-axis statistics
             ctx  = P^T V  per head            [B, 4, 32, 32]   pixel-axis reduction (existing linattn_ctx kernel)
             Out  = conv1x1(Q;   Wt[(h,e)][(h,d)] = ctx_h[d][e])
  backward   dQ   = conv1x1(dOut; Wt[(h,d)][(h,e)] = ctx_h[d][e])
             dctx = Q^T dOut per head                           pixel-axis reduction (existing linattn_bwd_dctx kernel)
             dV   = conv1x1(P;   Wt[(h,e)][(h,d)] = dctx_h[d][e])
             T    = conv1x1(V;   Wt[(h,d)][(h,e)] = dctx_h[d][e])
             dK   = P * (T - c),  c[(h,d)] = sum_e dctx_h[d][e] ctx_h[d][e]        elementwise
"""
import torch

HEADS, D = 4, 32


def blockdiag(m, transpose):
    """m: [B, HEADS, D, D] indexed [d][e].  transpose=False -> Wt[(h,e)][(h,d)] = m[d][e]; True -> Wt[(h,d)][(h,e)] = m[d][e]."""
    B = m.shape[0]
    w = torch.zeros(B, HEADS * D, HEADS * D, dtype=m.dtype)
    for h in range(HEADS):
        blk = m[:, h]                                   # [B, d, e]
        w[:, h * D:(h + 1) * D, h * D:(h + 1) * D] = blk if transpose else blk.transpose(1, 2)
    return w


def conv1x1_img(x, wt):
    """x: [B, n, C_in], wt: [B, C_out, C_in] -> [B, n, C_out]."""
    return torch.einsum("bnk,bok->bno", x, wt)


def forward(q, k, v):
    B, n, _ = q.shape
    p = torch.softmax(k, dim=1)                                            # over pixels, per channel
    ph, vh = p.view(B, n, HEADS, D), v.view(B, n, HEADS, D)
    ctx = torch.einsum("bnhd,bnhe->bhde", ph, vh)
    out = conv1x1_img(q, blockdiag(ctx, transpose=False))
    return out, (p, ctx)


def backward(q, v, saved, d_out):
    p, ctx = saved
    B, n, _ = q.shape
    dq = conv1x1_img(d_out, blockdiag(ctx, transpose=True))
    dctx = torch.einsum("bnhd,bnhe->bhde", q.view(B, n, HEADS, D), d_out.view(B, n, HEADS, D))
    dv = conv1x1_img(p, blockdiag(dctx, transpose=False))
    t = conv1x1_img(v, blockdiag(dctx, transpose=True))
    c = (dctx * ctx).sum(dim=3).reshape(B, 1, HEADS * D)
    dk = p * (t - c)
    return dq, dk, dv
