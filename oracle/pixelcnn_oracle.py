"""CPU restatement of the reference PixelCNN (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Follows /root/reference/src/models/pixelcnn.py:
    MaskedConvolution :12-24, VerticalStackConvolution :27-33, HorizontalStackConvolution :36-42,
    GatedMaskedConv :45-82 (vertical gate tanh*sigmoid :69, horizontal gate tanh*tanh :77),
    PixelCNN.__init__ :86-126, forward :128-154, calc_likelihood :156-165, sample :167-195.

``params`` uses the reference state_dict keys.  ``sample`` reproduces the reference loop (full forward
on the top h+1 rows for every pixel) but draws with an INJECTED uniform per (pixel, sample, channel)
through the inverse CDF  k = #{j : cdf_j <= u}  (or greedy argmax), because torch.multinomial's
stream cannot be shared with a CUDA kernel; the tests patch the reference's torch.multinomial with
the same rule.

PARITY PINNING: bit-exact against the unmodified reference on CPU (tests/test_pixelcnn.py, where
/root/reference is mounted) and against tests/golden/pixelcnn_*.npz (tests/golden/make_golden_pixelcnn.py).
"""
from collections import OrderedDict
from typing import Dict, Optional

import torch
import torch.nn.functional as F

DILATIONS = (1, 2, 1, 4, 1, 2, 1, 4, 1, 2, 1)   # pixelcnn.py:108-122


def v_mask(k: int, mask_center: bool) -> torch.Tensor:
    m = torch.ones(k, k)
    m[k // 2 + 1:, :] = 0
    if mask_center:
        m[k // 2] = 0
    return m


def h_mask(k: int, mask_center: bool) -> torch.Tensor:
    m = torch.ones(1, k)
    m[0, k // 2 + 1:] = 0
    if mask_center:
        m[0, k // 2] = 0
    return m


def param_shapes(channels: int, hidden: int, n_classes: Optional[int] = None) -> "OrderedDict[str, tuple]":
    """Ordered name -> shape identical to reference PixelCNN(...).state_dict(); n_classes: class_condition=True."""
    o = OrderedDict()
    o["log2"] = ()
    o["conv_vstack.mask"] = (5, 5)
    o["conv_vstack.conv.weight"] = (hidden, channels, 5, 5)
    o["conv_vstack.conv.bias"] = (hidden,)
    o["conv_hstack.mask"] = (1, 5)
    o["conv_hstack.conv.weight"] = (hidden, channels, 1, 5)
    o["conv_hstack.conv.bias"] = (hidden,)
    for i in range(len(DILATIONS)):
        p = f"conv_layers.{i}"
        o[f"{p}.horiz_conv.mask"] = (1, 3)
        o[f"{p}.horiz_conv.conv.weight"] = (2 * hidden, hidden, 1, 3)
        o[f"{p}.horiz_conv.conv.bias"] = (2 * hidden,)
        o[f"{p}.vert_conv.mask"] = (3, 3)
        o[f"{p}.vert_conv.conv.weight"] = (2 * hidden, hidden, 3, 3)
        o[f"{p}.vert_conv.conv.bias"] = (2 * hidden,)
        o[f"{p}.conv1x1_1.weight"] = (2 * hidden, 2 * hidden, 1, 1)
        o[f"{p}.conv1x1_1.bias"] = (2 * hidden,)
        o[f"{p}.conv1x1_2.weight"] = (hidden, hidden, 1, 1)
        o[f"{p}.conv1x1_2.bias"] = (hidden,)
        if n_classes is not None:      # :58-62
            for c in ("vert1", "vert2", "horiz1", "horiz2"):
                o[f"{p}.cond_proj_{c}.weight"] = (hidden, n_classes, 1, 1)
    o["conv_out.weight"] = (256 * channels, hidden, 1, 1)
    o["conv_out.bias"] = (256 * channels,)
    return o


def init_params(channels: int, hidden: int, seed: int = 0, n_classes: Optional[int] = None) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for name, shape in param_shapes(channels, hidden, n_classes).items():
        if name == "log2":
            out[name] = torch.log(torch.tensor(2.0))
        elif name.endswith("mask"):
            k = shape[1]
            center = name.startswith("conv_vstack") or name.startswith("conv_hstack")
            out[name] = v_mask(k, center) if shape[0] > 1 else h_mask(k, center)
        elif name.endswith("weight"):
            fan_in = shape[1] * shape[2] * shape[3]
            out[name] = (torch.rand(shape, generator=g) * 2 - 1) / fan_in ** 0.5
        else:
            out[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
    return out


def _masked(p, name, x, dilation=1):
    """MaskedConvolution.forward (:22-24): the weight is masked IN PLACE (``weight.data *= mask``), so autograd
    sees the masked weight as the leaf and masked taps still receive a gradient; padding = dilation*(k-1)//2 (:18)."""
    w = p[f"{name}.conv.weight"]
    w.data *= p[f"{name}.mask"]
    kh, kw = w.shape[2], w.shape[3]
    pad = (dilation * (kh - 1) // 2, dilation * (kw - 1) // 2)
    return F.conv2d(x, w, p[f"{name}.conv.bias"], padding=pad, dilation=dilation)


def forward(p: Dict[str, torch.Tensor], x: torch.Tensor, y: Optional[torch.Tensor] = None) -> torch.Tensor:
    """PixelCNN.forward (:128-154) -> logits [N, 256, C, H, W]; y = one-hot labels [N, n_classes] or None."""
    v = _masked(p, "conv_vstack", x)
    h = _masked(p, "conv_hstack", x)
    if y is not None:
        y = y.reshape(x.shape[0], -1, 1, 1)
    for i, d in enumerate(DILATIONS):
        n = f"conv_layers.{i}"
        vc = _masked(p, f"{n}.vert_conv", v, d)
        v1, v2 = torch.chunk(vc, 2, dim=1)
        if y is None:
            v_out = torch.tanh(v1) * torch.sigmoid(v2)
        else:                                            # :71
            v_out = torch.tanh(v1 + F.conv2d(y, p[f"{n}.cond_proj_vert1.weight"]).expand_as(v1)) * torch.sigmoid(
                v2 + F.conv2d(y, p[f"{n}.cond_proj_vert2.weight"]).expand_as(v2))
        hc = _masked(p, f"{n}.horiz_conv", h, d) + F.conv2d(vc, p[f"{n}.conv1x1_1.weight"], p[f"{n}.conv1x1_1.bias"])
        h1, h2 = torch.chunk(hc, 2, 1)
        if y is None:
            h_out = torch.tanh(h1) * torch.tanh(h2)      # sic: tanh*tanh on the horizontal stack (:77)
        else:                                            # :79
            h_out = torch.tanh(h1 + F.conv2d(y, p[f"{n}.cond_proj_horiz1.weight"]).expand_as(h1)) * torch.tanh(
                h2 + F.conv2d(y, p[f"{n}.cond_proj_horiz2.weight"]).expand_as(h2))
        h = F.conv2d(h_out, p[f"{n}.conv1x1_2.weight"], p[f"{n}.conv1x1_2.bias"]) + h
        v = v_out
    out = F.conv2d(F.elu(h), p["conv_out.weight"], p["conv_out.bias"])
    return out.reshape(out.shape[0], 256, out.shape[1] // 256, out.shape[2], out.shape[3])


def calc_likelihood(p, x, input_normalize: bool, y: Optional[torch.Tensor] = None):
    """:156-165 — bits per dimension."""
    pred = forward(p, x, y)
    target = ((x + 1) / 2 * 255).to(torch.long) if input_normalize else (x * 255).to(torch.long)
    nll = F.cross_entropy(pred, target, reduction="none")
    return (nll.mean(dim=[1, 2, 3]) / p["log2"]).mean()


def pick(probs: torch.Tensor, u: Optional[torch.Tensor]) -> torch.Tensor:
    """Inverse-CDF draw k = #{j : cdf_j <= u} (clamped to 255) or, with u None, the argmax."""
    if u is None:
        return probs.argmax(dim=-1)
    cdf = torch.cumsum(probs, dim=-1)
    return (cdf <= u[:, None]).sum(dim=-1).clamp(max=probs.shape[-1] - 1)


@torch.no_grad()
def sample(p, img_shape, uniforms: Optional[torch.Tensor], input_normalize: bool = False, img=None, y=None):
    """PixelCNN.sample (:167-195) with injected uniforms [H*W, N*C] (None = greedy)."""
    N, C, H, W = img_shape
    if img is None:
        img = torch.zeros(img_shape, dtype=torch.float32) - 1
    for h in range(H):
        for w in range(W):
            if (img[:, :, h, w] != -1).all().item():
                continue
            pred = forward(p, img[:, :, : h + 1, :], y)
            probs = F.softmax(pred[:, :, :, h, w].permute(0, 2, 1), dim=-1).reshape(N * C, -1)
            k = pick(probs, None if uniforms is None else uniforms[h * W + w])
            new = k.to(torch.float32) / 255
            if input_normalize:
                new = new * 2 - 1
            img[:, :, h, w] = new.reshape(N, C)
    return img
