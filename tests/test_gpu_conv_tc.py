"""GPU: the tcgen05 (bf16x3) convolution kernel in isolation, through the C ABI's test entry
igm_debug_conv, against torch's fp64 CPU convolution (the oracle for a single conv) and against
the SIMT fp32 engine.  Tolerance 1e-4 relative: bf16x3 carries ~16 mantissa bits per operand."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from igm_b200 import _lib
from tests._util import rel_err

pytestmark = pytest.mark.gpu

SHAPES = [
    # B, H, W, Cin, Cout, K
    (2, 32, 32, 64, 64, 3),
    (2, 32, 32, 64, 128, 3),
    (2, 32, 32, 128, 64, 3),
    (2, 16, 16, 128, 128, 3),
    (4, 8, 8, 256, 256, 3),
    (3, 8, 8, 512, 128, 3),      # two images per tile, odd batch
    (1, 8, 8, 256, 256, 3),      # batch smaller than the tile's image count
    (2, 32, 32, 64, 384, 1),     # to_qkv
    (2, 16, 16, 128, 64, 1),
    (2, 28, 28, 128, 128, 3),    # ragged: 112-row tiles
    (2, 14, 14, 256, 256, 3),    # ragged: 9-row tiles over 14 rows
    (1, 64, 64, 64, 64, 3),      # CelebA resolution: 2-row tiles
    (5, 4, 4, 64, 64, 3),        # 8 images per tile
]


def _run(engine, mode, x, w, bias, add, B, H, W, Cin, Cout, K):
    lib = _lib.load()
    N = Cout if mode == 0 else Cin
    out = torch.empty(B, H, W, N, device="cuda", dtype=torch.float32)
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    rc = lib.igm_debug_conv(engine, mode, p(x), p(w), p(bias), p(add), p(out), B, H, W, Cin, Cout, K, None)
    assert rc == 0, lib.igm_last_error(None).decode()
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("shape", SHAPES)
def test_conv_tc_forward(shape):
    B, H, W, Cin, Cout, K = shape
    g = torch.Generator().manual_seed(hash(shape) % 1000)
    x = torch.randn(B, H, W, Cin, generator=g)
    w = torch.randn(Cout, Cin, K, K, generator=g) / (Cin * K * K) ** 0.5
    bias = torch.randn(Cout, generator=g)
    add = torch.randn(B, H, W, Cout, generator=g)
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), bias.double(), padding=(K - 1) // 2)
    ref = ref.permute(0, 2, 3, 1) + add.double()
    xs, ws, bs, ads = x.cuda(), w.cuda(), bias.cuda(), add.cuda()
    tc = _run(1, 0, xs, ws, bs, ads, B, H, W, Cin, Cout, K).cpu()
    simt = _run(0, 0, xs, ws, bs, ads, B, H, W, Cin, Cout, K).cpu()
    l2, mx = rel_err(tc, ref)
    l2s, mxs = rel_err(simt, ref)
    assert l2s < 1e-5 and mxs < 1e-5, f"SIMT engine off: {l2s:.2e} {mxs:.2e}"
    assert l2 < 1e-4 and mx < 1e-4, f"tcgen05 engine: rel-L2 {l2:.2e} max-rel {mx:.2e}"


@pytest.mark.parametrize("shape", SHAPES[:8])
def test_conv_tc_dgrad(shape):
    B, H, W, Cin, Cout, K = shape
    g = torch.Generator().manual_seed(1 + hash(shape) % 1000)
    dy = torch.randn(B, H, W, Cout, generator=g)
    w = torch.randn(Cout, Cin, K, K, generator=g) / (Cout * K * K) ** 0.5
    ref = F.conv_transpose2d(dy.permute(0, 3, 1, 2).double(), w.double(), padding=(K - 1) // 2).permute(0, 2, 3, 1)
    tc = _run(1, 1, dy.cuda(), w.cuda(), None, None, B, H, W, Cin, Cout, K).cpu()
    simt = _run(0, 1, dy.cuda(), w.cuda(), None, None, B, H, W, Cin, Cout, K).cpu()
    l2s, mxs = rel_err(simt, ref)
    assert l2s < 1e-5 and mxs < 1e-5, f"SIMT engine off: {l2s:.2e} {mxs:.2e}"
    l2, mx = rel_err(tc, ref)
    assert l2 < 1e-4 and mx < 1e-4, f"tcgen05 engine: rel-L2 {l2:.2e} max-rel {mx:.2e}"


WSHAPES = [
    # B, H, W, Cin, Cout, K
    (2, 32, 32, 64, 64, 3),
    (2, 32, 32, 64, 128, 3),
    (2, 32, 32, 128, 64, 3),
    (4, 16, 16, 128, 128, 3),
    (4, 8, 8, 256, 256, 3),
    (3, 8, 8, 512, 128, 3),
    (2, 32, 32, 64, 384, 1),
    (2, 16, 16, 128, 64, 1),
    (2, 28, 28, 128, 128, 3),    # ragged 56-row boxes (zero-initialised smem rows)
    (2, 14, 14, 256, 256, 3),
    (1, 64, 64, 64, 64, 3),
    (8, 4, 4, 64, 64, 3),        # 4 images per box
]


def _wgrad(engine, variant, x, dy, B, H, W, Cin, Cout, K):
    lib = _lib.load()
    gw = torch.zeros(Cout, Cin, K, K, device="cuda", dtype=torch.float32)
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.igm_debug_wgrad(engine, variant, p(x), p(dy), p(gw), B, H, W, Cin, Cout, K, None)
    assert rc == 0, lib.igm_last_error(None).decode()
    torch.cuda.synchronize()
    return gw


@pytest.mark.parametrize("shape", WSHAPES)
def test_wgrad_tc(shape):
    B, H, W, Cin, Cout, K = shape
    g = torch.Generator().manual_seed(2 + hash(shape) % 1000)
    x = torch.randn(B, H, W, Cin, generator=g)
    dy = torch.randn(B, H, W, Cout, generator=g)
    xr = x.permute(0, 3, 1, 2).double()
    w = torch.zeros(Cout, Cin, K, K, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(xr, w, padding=(K - 1) // 2)
    ref, = torch.autograd.grad(y, w, dy.permute(0, 3, 1, 2).double())
    simt = _wgrad(0, 0, x.cuda(), dy.cuda(), B, H, W, Cin, Cout, K).cpu()
    l2s, mxs = rel_err(simt, ref)
    assert l2s < 1e-5 and mxs < 1e-5, f"SIMT wgrad off: {l2s:.2e} {mxs:.2e}"
    tc = _wgrad(1, 0, x.cuda(), dy.cuda(), B, H, W, Cin, Cout, K).cpu()
    l2, mx = rel_err(tc, ref)
    assert l2 < 1e-4 and mx < 1e-4, f"tcgen05 wgrad: rel-L2 {l2:.2e} max-rel {mx:.2e}"


RSHAPES = [
    # kind, B, H, W, C, C2     (H, W = INPUT size of the layer)
    (0, 2, 32, 32, 64, 64),      # Downsample 32 -> 16
    (0, 4, 16, 16, 128, 128),    # Downsample 16 -> 8 (two images per tile)
    (0, 2, 28, 28, 128, 128),    # ragged 14x14 output
    (0, 1, 64, 64, 64, 64),      # CelebA 64 -> 32
    (1, 4, 8, 8, 128, 128),      # Upsample 8 -> 16
    (1, 2, 16, 16, 64, 64),      # Upsample 16 -> 32
    (1, 2, 14, 14, 128, 128),    # ragged
    (1, 3, 8, 8, 256, 64),       # C != C2 (not in the U-Net, exercises the layout strides)
]


def _resample(engine, kind, mode, x, aux, w, bias, add, B, H, W, Cc, C2):
    lib = _lib.load()
    OH, OW = (H // 2, W // 2) if kind == 0 else (2 * H, 2 * W)
    if mode == 0:
        out = torch.empty(B, OH, OW, C2, device="cuda")
    elif mode == 1:
        out = torch.empty(B, H, W, Cc, device="cuda")
    else:
        out = torch.zeros_like(w)
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    rc = lib.igm_debug_resample(engine, kind, mode, p(x), p(aux), p(w), p(bias), p(add), p(out), B, H, W, Cc, C2, None)
    assert rc == 0, lib.igm_last_error(None).decode()
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("shape", RSHAPES)
def test_resample_tc(shape):
    kind, B, H, W, Cc, C2 = shape
    K = 3 if kind == 0 else 4
    OH, OW = (H // 2, W // 2) if kind == 0 else (2 * H, 2 * W)
    g = torch.Generator().manual_seed(3 + hash(shape) % 1000)
    x = torch.randn(B, H, W, Cc, generator=g)
    dy = torch.randn(B, OH, OW, C2, generator=g)
    wshape = (C2, Cc, K, K) if kind == 0 else (Cc, C2, K, K)
    w = torch.randn(wshape, generator=g) / (Cc * K * K / (1 if kind == 0 else 4)) ** 0.5
    bias = torch.randn(C2, generator=g)
    add = torch.randn(B, H, W, Cc, generator=g)
    xr = x.permute(0, 3, 1, 2).double().requires_grad_(True)
    wr = w.double().requires_grad_(True)
    if kind == 0:
        y = F.conv2d(xr, wr, bias.double(), stride=2, padding=1)
    else:
        y = F.conv_transpose2d(xr, wr, bias.double(), stride=2, padding=1)
    gx, gw = torch.autograd.grad(y, [xr, wr], dy.permute(0, 3, 1, 2).double())
    ref = {0: y.detach().permute(0, 2, 3, 1), 1: gx.permute(0, 2, 3, 1) + add.double(), 2: gw}
    xs, dys, ws, bs, ads = x.cuda(), dy.cuda(), w.cuda(), bias.cuda(), add.cuda()
    for mode in (0, 1, 2):
        args = {0: (xs, None, ws, bs, None), 1: (dys, None, ws, None, ads), 2: (xs, dys, ws, None, None)}[mode]
        simt = _resample(0, kind, mode, *args, B, H, W, Cc, C2).cpu()
        l2s, mxs = rel_err(simt, ref[mode])
        assert l2s < 1e-5 and mxs < 1e-5, f"SIMT kind {kind} mode {mode}: {l2s:.2e} {mxs:.2e}"
        tc = _resample(1, kind, mode, *args, B, H, W, Cc, C2).cpu()
        l2, mx = rel_err(tc, ref[mode])
        assert l2 < 1e-4 and mx < 1e-4, f"tcgen05 kind {kind} mode {mode}: rel-L2 {l2:.2e} max-rel {mx:.2e}"


import os

# ---- CTA-pair (cta_group::2) engine (csrc/conv_tc2.cu): parity-green on B200 (round 2) but slower than the per-tap
# engine, so it is not on the product path; the kernel-level tests run by default, the whole-net one on request ----
_pair = pytest.mark.skipif(False, reason="")
_pair_net = pytest.mark.skipif(os.environ.get("IGM_TEST_CONV_PAIR") != "1",
                               reason="whole-net parity with the (slower) CTA-pair engine: set IGM_TEST_CONV_PAIR=1")

PAIR_SHAPES = [
    # B, H, W, Cin, Cout, K   (Cout % 128 == 0 for fprop; dgrad needs Cin % 128 == 0)
    (1, 8, 32, 64, 128, 1),      # one pair tile, one tap, one K chunk
    (2, 16, 16, 128, 128, 3),    # 4 M tiles = 2 pairs
    (4, 8, 8, 256, 256, 3),      # two images per M tile, two N tiles
    (5, 8, 8, 128, 128, 3),      # 3 M tiles (two images each): the last pair's second tile does not exist
    (2, 32, 32, 128, 128, 3),
    (160, 16, 16, 128, 128, 3),  # more pair tiles than CTA pairs: persistent loop, both accumulator stages
]


@_pair
@pytest.mark.parametrize("shape", PAIR_SHAPES)
def test_conv_pair_forward(shape):
    B, H, W, Cin, Cout, K = shape
    g = torch.Generator().manual_seed(13 + hash(shape) % 1000)
    x = torch.randn(B, H, W, Cin, generator=g)
    w = torch.randn(Cout, Cin, K, K, generator=g) / (Cin * K * K) ** 0.5
    bias = torch.randn(Cout, generator=g)
    add = torch.randn(B, H, W, Cout, generator=g)
    xs, ws, bs, ads = x.cuda(), w.cuda(), bias.cuda(), add.cuda()
    ref = _run(1, 0, xs, ws, bs, ads, B, H, W, Cin, Cout, K).cpu()
    got = _run(3, 0, xs, ws, bs, ads, B, H, W, Cin, Cout, K).cpu()
    assert torch.isfinite(got).all()
    l2, mx = rel_err(got, ref.double())
    assert l2 < 1e-5 and mx < 1e-5, f"pair engine vs per-tap engine: rel-L2 {l2:.2e} max-rel {mx:.2e}"


@_pair
@pytest.mark.parametrize("shape", PAIR_SHAPES[1:5])
def test_conv_pair_dgrad(shape):
    B, H, W, Cin, Cout, K = shape
    g = torch.Generator().manual_seed(17 + hash(shape) % 1000)
    dy = torch.randn(B, H, W, Cout, generator=g)
    w = torch.randn(Cout, Cin, K, K, generator=g) / (Cout * K * K) ** 0.5
    ref = F.conv_transpose2d(dy.permute(0, 3, 1, 2).double(), w.double(), padding=(K - 1) // 2).permute(0, 2, 3, 1)
    got = _run(3, 1, dy.cuda(), w.cuda(), None, None, B, H, W, Cin, Cout, K).cpu()
    l2, mx = rel_err(got, ref)
    assert l2 < 1e-4 and mx < 1e-4, f"pair engine dgrad: rel-L2 {l2:.2e} max-rel {mx:.2e}"


@_pair_net
def test_unet_parity_with_pair_engine():
    """The whole parity suite with IGM_CONV_PAIR=1 (read once per process, hence the subprocess): every stride-1 conv
    with N % 128 == 0 then runs forward and data gradient on conv_tc2.cu, fused GroupNorm statistics included."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, IGM_CONV_PAIR="1")
    env.pop("IGM_TEST_CONV_PAIR", None)
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-x", "-q", "-m", "gpu"], cwd=root, env=env,
                       capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
