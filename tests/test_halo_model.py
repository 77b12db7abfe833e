"""CPU: the padded-linear tile -> pixel map planned for the halo-reuse forward conv (tools/halo_fprop_model.py) against
torch's conv2d, including ragged bands (H not a multiple of BH) and the 28x28 / 8x8 geometries."""
import importlib.util
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

_p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "halo_fprop_model.py")
_s = importlib.util.spec_from_file_location("halo_fprop_model", _p)
M = importlib.util.module_from_spec(_s)
_s.loader.exec_module(M)


@pytest.mark.parametrize("H,W,BH", [(8, 8, 3), (8, 8, 8), (16, 16, 7), (28, 28, 5), (32, 32, 7)])
def test_padded_linear_conv_equals_conv2d(H, W, BH):
    g = torch.Generator().manual_seed(H * 100 + BH)
    x = torch.randn(2, 5, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(6, 5, 3, 3, generator=g, dtype=torch.float64)
    ref = F.conv2d(x, w, padding=1).permute(0, 2, 3, 1).numpy()
    got = M.conv3x3_halo(x.permute(0, 2, 3, 1).numpy(), w.numpy(), BH)
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)


def test_band_plan_numbers():
    m, n, f, px = M.band_plan(32, 32, 7)
    assert (m, n) == (238, 2) and abs(f - 0.875) < 1e-9 and px == 9 * 34 + 2 * 34


@pytest.mark.parametrize("H,W", [(32, 32), (28, 28), (64, 64), (5, 32), (7, 40)])
def test_kernel_decomposition_and_row_stores(H, W):
    """csrc/conv_halo.cu's band / tile / warp decomposition and its clipped per-image-row stores: every output pixel
    is stored exactly once, from the right accumulator row, junk rows (NaN in the model) never reach the output, and
    the per-warp GroupNorm slots partition the valid pixels."""
    g = torch.Generator().manual_seed(H * 1000 + W)
    x = torch.randn(2, 4, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(3, 4, 3, 3, generator=g, dtype=torch.float64)
    bias = torch.randn(3, generator=g, dtype=torch.float64)
    ref = F.conv2d(x, w, bias, padding=1).permute(0, 2, 3, 1).numpy()
    out, writes, slots = M.conv3x3_halo_kernel_model(x.permute(0, 2, 3, 1).numpy(), w.numpy(), bias.numpy())
    assert (writes == 1).all()
    np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(slots.sum(axis=1), ref.sum(axis=(1, 2, 3)), rtol=1e-10, atol=1e-10)


def test_pair_kernel_column_algebra():
    """csrc/conv_tc2.cu: per-CTA weight staging [w_hi half ; w_lo half], the 64-column offset of the a_lo x w_hi MMA and
    the epilogue's column map reproduce a_hi*w_hi + a_hi*w_lo + a_lo*w_hi for every output channel."""
    _q = os.path.join(os.path.dirname(_p), "pair_columns_model.py")
    _t = importlib.util.spec_from_file_location("pair_columns_model", _q)
    PM = importlib.util.module_from_spec(_t)
    _t.loader.exec_module(PM)
    rng = np.random.default_rng(0)
    a_hi, a_lo = rng.standard_normal((256, 64)), 1e-3 * rng.standard_normal((256, 64))
    w_hi, w_lo = rng.standard_normal((128, 64)), 1e-3 * rng.standard_normal((128, 64))
    want = a_hi @ w_hi.T + a_hi @ w_lo.T + a_lo @ w_hi.T
    np.testing.assert_allclose(PM.pair_tile(a_hi, a_lo, w_hi, w_lo), want, rtol=1e-12, atol=1e-12)
