"""PixelCNN: oracle vs reference/golden on CPU; the incremental CUDA engine vs the oracle on the GPU.
Bars: logits 1e-3 relative; greedy-decoded pixels bit-exact; inverse-CDF pixels bit-exact except where
the draw is decided by less than 1e-5 of probability mass (checked against the oracle's own CDF)."""
import importlib.util
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import pixelcnn_oracle as PO
from oracle import ref_loader
from tests._util import GOLDEN, assert_close

_spec = importlib.util.spec_from_file_location("make_golden_pixelcnn", os.path.join(GOLDEN, "make_golden_pixelcnn.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)


def _golden(case):
    return dict(np.load(os.path.join(GOLDEN, f"pixelcnn_{case}.npz")))


@pytest.mark.parametrize("case", list(mg.CASES))
def test_oracle_matches_golden(case):
    C, Hd, N, H, W, norm = mg.CASES[case]
    p = PO.init_params(C, Hd, seed=1)
    x, u = mg.inputs(case)
    g = _golden(case)
    with torch.no_grad():
        logits = PO.forward(p, x)
    assert_close(logits[:, :, :, ::3, ::3], g["logits_sub"], "logits", 1e-5)
    assert abs(PO.calc_likelihood(p, x, norm).item() - g["bpd"]) < 1e-4
    sh, sw = mg.SAMPLE_HW[case]
    if case == "rgb_small":   # the MNIST crop costs ~1 min of CPU: covered by the fixture + GPU test instead
        assert np.array_equal(PO.sample(p, (N, C, sh, sw), u, input_normalize=norm).numpy(), g["sample_u"])
        assert np.array_equal(PO.sample(p, (N, C, sh, sw), None, input_normalize=norm).numpy(), g["sample_greedy"])


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_oracle_bit_exact_vs_live_reference():
    ref = ref_loader.load("pixelcnn")
    torch.manual_seed(0)
    m = ref.PixelCNN(ref_loader.datamodule_cfg(1, 12, 12, normalize=False), hidden_dim=32)
    sd = m.state_dict()
    shapes = PO.param_shapes(1, 32)
    assert list(sd.keys()) == list(shapes.keys()) and all(tuple(sd[k].shape) == shapes[k] for k in sd)
    x = torch.rand(2, 1, 12, 12)
    with torch.no_grad():
        assert torch.equal(m(x), PO.forward(sd, x))
        assert torch.equal(m.calc_likelihood(x), PO.calc_likelihood(sd, x, False))


def test_mirror_state_dict_and_packing():
    import igm_b200
    from igm_b200 import _lib
    dm = ref_loader.datamodule_cfg(1, 28, 28, normalize=False)
    torch.manual_seed(0)
    m = igm_b200.PixelCNN(dm, hidden_dim=64)
    shapes = PO.param_shapes(1, 64)
    sd = m.state_dict()
    assert list(sd.keys()) == list(shapes.keys()) and all(tuple(sd[k].shape) == shapes[k] for k in sd)
    flat = m._pack()
    assert flat.numel() == _lib.load().igm_pixelcnn_weight_floats(1, 64)
    if ref_loader.available():
        ref = ref_loader.load("pixelcnn")
        torch.manual_seed(0)
        r = ref.PixelCNN(dm, hidden_dim=64)
        assert all(torch.equal(r.state_dict()[k], sd[k]) for k in sd)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 1, 28, 28))


def _mirror(case):
    import igm_b200
    C, Hd, N, H, W, norm = mg.CASES[case]
    p = PO.init_params(C, Hd, seed=1)
    m = igm_b200.PixelCNN(ref_loader.datamodule_cfg(C, H, W, normalize=norm), hidden_dim=Hd)
    m.load_state_dict(p)
    return p, m.cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(mg.CASES))
def test_gpu_forward_logits(case):
    C, Hd, N, H, W, norm = mg.CASES[case]
    p, m = _mirror(case)
    x, u = mg.inputs(case)
    with torch.no_grad():
        ref = PO.forward(p, x)
        got = m(x.cuda()).cpu()
    assert_close(got, ref, f"{case} logits")
    assert_close(got[:, :, :, ::3, ::3], _golden(case)["logits_sub"], f"{case} logits vs reference fixture")
    assert abs(m.calc_likelihood(x.cuda()).item() - _golden(case)["bpd"]) < 1e-3 * _golden(case)["bpd"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(mg.CASES))
def test_gpu_sampler_pixels(case):
    C, Hd, N, H, W, norm = mg.CASES[case]
    p, m = _mirror(case)
    x, u = mg.inputs(case)
    g = _golden(case)
    sh, sw = mg.SAMPLE_HW[case]
    # greedy decode: bit-exact pixels
    s_g = m.sample((N, C, sh, sw), greedy=True).cpu()
    assert np.array_equal(s_g.numpy(), g["sample_greedy"]), "greedy pixels differ from the reference fixture"
    # inverse-CDF with shared uniforms: bit-exact, or decided by < 1e-5 of probability mass
    s_u = m.sample((N, C, sh, sw), uniforms=u.cuda()).cpu()
    if not np.array_equal(s_u.numpy(), g["sample_u"]):
        with torch.no_grad():
            logits = PO.forward(p, s_u)          # teacher-forced oracle on the GPU's own pixels
        probs = F.softmax(logits, dim=1)         # [N, 256, C, h, w]
        cdf = torch.cumsum(probs, dim=1)
        k = ((s_u + 1) / 2 * 255 if norm else s_u * 255).round().long()
        for hh in range(sh):
            for ww in range(sw):
                uu = u[hh * sw + ww].reshape(N, C)
                c = cdf[:, :, :, hh, ww]
                kk = k[:, :, hh, ww]
                lo = torch.where(kk > 0, c.gather(1, (kk - 1).clamp(min=0)[:, None])[:, 0], torch.zeros_like(uu))
                hi = c.gather(1, kk[:, None])[:, 0]
                assert bool(((lo - 1e-5 <= uu) & ((uu < hi + 1e-5) | (kk == 255))).all()), f"pixel ({hh},{ww}) is not a valid draw"
    # given pixels are kept, -1 pixels are generated (reference :179-186)
    start = torch.full((N, C, sh, sw), -1.0)
    start[:, :, : sh // 2, :] = torch.from_numpy(g["sample_greedy"])[:, :, : sh // 2, :]
    cont = m.sample((N, C, sh, sw), img=start.clone(), greedy=True).cpu()
    assert np.array_equal(cont.numpy(), g["sample_greedy"])


@pytest.mark.gpu
def test_gpu_full_size_sample_is_causally_consistent():
    """BASELINE config C4 (MNIST 28x28, hidden 64, batch 64): every drawn pixel is a valid inverse-CDF
    draw under the engine's own teacher-forced logits, and Philox draws are seeded."""
    import igm_b200
    torch.manual_seed(0)
    m = igm_b200.PixelCNN(ref_loader.datamodule_cfg(1, 28, 28, normalize=False), hidden_dim=64).cuda()
    g = torch.Generator().manual_seed(4)
    u = torch.rand(784, 64, generator=g).cuda()
    img = m.sample((64, 1, 28, 28), uniforms=u)
    assert float(img.min()) >= 0 and float(img.max()) <= 1
    logits = m(img)
    cdf = torch.cumsum(F.softmax(logits, dim=1), dim=1)[:, :, 0]          # [64, 256, 28, 28]
    k = (img[:, 0] * 255).round().long()
    uu = u.reshape(28, 28, 64).permute(2, 0, 1)
    hi = cdf.gather(1, k[:, None])[:, 0]
    lo = torch.where(k > 0, cdf.gather(1, (k - 1).clamp(min=0)[:, None])[:, 0], torch.zeros_like(hi))
    ok = (lo - 1e-5 <= uu) & ((uu < hi + 1e-5) | (k == 255))
    assert bool(ok.all()), f"{int((~ok).sum())} pixels are not valid draws"
    a = m.sample((4, 1, 28, 28), seed=7)
    b = m.sample((4, 1, 28, 28), seed=7)
    c = m.sample((4, 1, 28, 28), seed=8)
    assert torch.equal(a, b) and not torch.equal(a, c)
