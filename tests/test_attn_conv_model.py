"""CPU: the per-image 1x1-convolution formulation of linear attention planned for the training path
(tools/attn_conv_model.py) against autograd of the reference formulation (src/models/ddpm.py:154-166)."""
import importlib.util
import os

import torch

_p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "attn_conv_model.py")
_s = importlib.util.spec_from_file_location("attn_conv_model", _p)
A = importlib.util.module_from_spec(_s)
_s.loader.exec_module(A)


def _reference(qkv_nchw):
    """ddpm.py:157-163 on the to_qkv output [B, 384, H, W] (rearrange spelled with reshape)."""
    b, _, h, w = qkv_nchw.shape
    qkv = qkv_nchw.reshape(b, 3, 4, 32, h * w)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    k = k.softmax(dim=-1)
    context = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", context, q)
    return out.reshape(b, 128, h, w)


def test_conv_formulation_matches_reference_autograd():
    torch.manual_seed(0)
    B, H, W = 2, 4, 6
    qkv = torch.randn(B, 384, H, W, dtype=torch.float64, requires_grad=True)
    g = torch.randn(B, 128, H, W, dtype=torch.float64)
    ref = _reference(qkv)
    (d_qkv,) = torch.autograd.grad(ref, qkv, g)
    nhwc = lambda t: t.permute(0, 2, 3, 1).reshape(B, H * W, -1)
    x = nhwc(qkv.detach())
    q, k, v = x[..., :128], x[..., 128:256], x[..., 256:]
    out, saved = A.forward(q, k, v)
    torch.testing.assert_close(out, nhwc(ref.detach()), rtol=1e-12, atol=1e-12)
    dq, dk, dv = A.backward(q, v, saved, nhwc(g))
    want = nhwc(d_qkv)
    torch.testing.assert_close(dq, want[..., :128], rtol=1e-11, atol=1e-12)
    torch.testing.assert_close(dk, want[..., 128:256], rtol=1e-11, atol=1e-12)
    torch.testing.assert_close(dv, want[..., 256:], rtol=1e-11, atol=1e-12)
