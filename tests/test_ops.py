"""Generic CUDA operators (csrc/ops.cu through igm_b200.ops) vs torch's CPU fp32 reference of the same op:
every convolution geometry the VQ-VAE networks (src/networks/vqvae.py) and PixelCNN (src/models/pixelcnn.py)
use, the activations / gates, the 256-way cross entropy, MSE and the straight-through value.  Bar: 1e-3 rel."""
import pytest
import torch
import torch.nn.functional as F

from tests._util import assert_close

CONVS = [
    # (B, Cin, H, W, Cout, KH, KW, stride, pad, dil, transposed, bias)
    (2, 3, 16, 16, 8, 4, 4, 2, (1, 1), 1, False, True),      # encoder stem k4 s2 p1
    (2, 8, 8, 8, 16, 3, 3, 1, (1, 1), 1, False, False),      # residual 3x3 without bias
    (2, 16, 8, 8, 8, 1, 1, 1, (0, 0), 1, False, False),      # residual 1x1
    (2, 16, 4, 4, 32, 3, 3, 1, (1, 1), 1, True, True),       # decoder ConvTranspose k3 s1 p1
    (2, 32, 4, 4, 16, 4, 4, 2, (1, 1), 1, True, True),       # decoder ConvTranspose k4 s2 p1
    (2, 16, 8, 8, 3, 4, 4, 2, (1, 1), 1, True, True),
    (2, 1, 9, 11, 32, 5, 5, 1, (2, 2), 1, False, True),      # PixelCNN vertical stem
    (2, 3, 9, 11, 32, 1, 5, 1, (0, 2), 1, False, True),      # PixelCNN horizontal stem (rectangular)
    (2, 32, 9, 11, 64, 3, 3, 1, (2, 2), 2, False, True),     # dilated vertical
    (2, 32, 9, 11, 64, 1, 3, 1, (0, 4), 4, False, True),     # dilated horizontal (rectangular)
    (3, 4, 1, 1, 32, 1, 1, 1, (0, 0), 1, False, False),      # cond_proj on [N, n_classes, 1, 1]
    (1, 5, 7, 5, 6, 3, 3, 2, (1, 1), 1, False, True),        # odd sizes, stride 2
    # 64-multiple channels: the tensor-core route of the operator (csrc/ops_tc.cu, tcgen05 bf16x3)
    (2, 64, 8, 8, 128, 3, 3, 1, (1, 1), 1, False, True),     # VQ-VAE residual 3x3 (two 8x8 images per MMA tile)
    (2, 128, 8, 8, 64, 1, 1, 1, (0, 0), 1, False, False),    # VQ-VAE residual 1x1
    (3, 64, 32, 32, 64, 3, 3, 1, (1, 1), 1, False, True),    # encoder 3x3 at the C5 latent size, odd batch
    (2, 64, 8, 8, 128, 3, 3, 1, (1, 1), 1, True, True),      # decoder ConvTranspose k3 s1 p1 64 -> 128
    (2, 128, 8, 8, 64, 4, 4, 2, (1, 1), 1, True, True),      # decoder ConvTranspose k4 s2 p1 128 -> 64 (phases / strided plans)
    (3, 128, 16, 16, 64, 4, 4, 2, (1, 1), 1, True, False),
]


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", CONVS, ids=lambda c: "x".join(str(v) for v in c[:10]) + ("T" if c[10] else ""))
def test_conv2d_forward_backward(cfg):
    from igm_b200 import ops
    B, Ci, H, W, Co, KH, KW, s, pad, dil, tr, bias = cfg
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, Ci, H, W, generator=g, requires_grad=True)
    w = (torch.randn((Ci, Co, KH, KW) if tr else (Co, Ci, KH, KW), generator=g) * 0.2).requires_grad_(True)
    b = torch.randn(Co, generator=g).requires_grad_(True) if bias else None
    ref = F.conv_transpose2d(x, w, b, s, pad) if tr else F.conv2d(x, w, b, s, pad, dil)
    res = torch.randn(ref.shape, generator=g, requires_grad=True)
    dy = torch.randn(ref.shape, generator=g)
    (ref + res).backward(dy)
    xc, wc, rc = (t.detach().cuda().requires_grad_(True) for t in (x, w, res))
    bc = b.detach().cuda().requires_grad_(True) if bias else None
    got = ops.conv_transpose2d(xc, wc, bc, s, pad, residual=rc) if tr else ops.conv2d(xc, wc, bc, s, pad, dil, residual=rc)
    assert got.shape == ref.shape
    got.backward(dy.cuda())
    assert_close(got, ref + res, "y")
    assert_close(xc.grad, x.grad, "dx")
    assert_close(wc.grad, w.grad, "dw")
    assert_close(rc.grad, res.grad, "dresidual")
    if bias:
        assert_close(bc.grad, b.grad, "db")


THIN_S2 = [
    # (B, Cin, H, W, Cout, transposed): 4x4 stride-2 pad-1 layers with <= 4 channels on one side, no residual -- the first conv
    # and the last transposed conv of the VQ-VAE (csrc/conv_simt.cu conv_thin_in_s2 / convT_thin_out_s2 / wgrad_thin_s2 kernels)
    (2, 3, 16, 16, 32, False),     # encoder stem at hidden 32 (one output channel per lane)
    (3, 3, 32, 24, 64, False),     # two output channels per lane, odd batch, non-square
    (2, 1, 8, 8, 32, False),
    (2, 32, 8, 8, 3, True),        # decoder head
    (3, 64, 12, 16, 3, True),
    (2, 32, 4, 4, 1, True),
]


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", THIN_S2, ids=lambda c: "x".join(str(v) for v in c[:5]) + ("T" if c[5] else ""))
def test_thin_stride2_layers(cfg):
    from igm_b200 import ops
    B, Ci, H, W, Co, tr = cfg
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, Ci, H, W, generator=g, requires_grad=True)
    w = (torch.randn((Ci, Co, 4, 4) if tr else (Co, Ci, 4, 4), generator=g) * 0.2).requires_grad_(True)
    b = torch.randn(Co, generator=g).requires_grad_(True)
    ref = F.conv_transpose2d(x, w, b, 2, 1) if tr else F.conv2d(x, w, b, 2, 1)
    dy = torch.randn(ref.shape, generator=g)
    ref.backward(dy)
    xc, wc, bc = (t.detach().cuda().requires_grad_(True) for t in (x, w, b))
    got = ops.conv_transpose2d(xc, wc, bc, 2, (1, 1)) if tr else ops.conv2d(xc, wc, bc, 2, (1, 1), 1)
    assert got.shape == ref.shape
    got.backward(dy.cuda())
    assert_close(got, ref, "y")
    assert_close(xc.grad, x.grad, "dx")
    assert_close(wc.grad, w.grad, "dw")
    assert_close(bc.grad, b.grad, "db")


@pytest.mark.gpu
def test_activations_and_losses():
    from igm_b200 import ops
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, 8, 5, 7, generator=g)
    for name, fn, ref in (("relu", ops.relu, F.relu), ("elu", ops.elu, F.elu)):
        a = x.clone().requires_grad_(True)
        c = x.cuda().requires_grad_(True)
        dy = torch.randn(x.shape, generator=g)
        ref(a).backward(dy)
        y = fn(c)
        y.backward(dy.cuda())
        assert_close(y, ref(a), name)
        assert_close(c.grad, a.grad, name + " grad")
    cond = torch.randn(3, 8, generator=g)
    for name, fn, second in (("tanh*sigmoid", ops.gate_tanh_sigmoid, torch.sigmoid), ("tanh*tanh", ops.gate_tanh_tanh, torch.tanh)):
        for use_cond in (False, True):
            a = x.clone().requires_grad_(True)
            ca = cond.clone().requires_grad_(True)
            pre = a + ca[:, :, None, None] if use_cond else a
            a1, a2 = torch.chunk(pre, 2, dim=1)
            r = torch.tanh(a1) * second(a2)
            dy = torch.randn(r.shape, generator=g)
            r.backward(dy)
            c = x.cuda().requires_grad_(True)
            cc = cond.cuda().requires_grad_(True)
            y = fn(c, cc if use_cond else None)
            y.backward(dy.cuda())
            assert_close(y, r, name)
            assert_close(c.grad, a.grad, name + " grad")
            if use_cond:
                assert_close(cc.grad, ca.grad, name + " cond grad")
    # 256-way cross entropy on the conv_out layout [N, 256*C, H, W]
    N, Cc, H, W = 2, 3, 4, 5
    logits = torch.randn(N, 256 * Cc, H, W, generator=g).requires_grad_(True)
    target = torch.randint(0, 256, (N, Cc, H, W), generator=g)
    wgt = torch.rand(N, Cc, H, W, generator=g)
    ref = F.cross_entropy(logits.reshape(N, 256, Cc, H, W), target, reduction="none")
    (ref * wgt).sum().backward()
    lc = logits.detach().cuda().requires_grad_(True)
    got = ops.cross_entropy_256(lc, target.cuda())
    (got * wgt.cuda()).sum().backward()
    assert_close(got, ref, "nll")
    assert_close(lc.grad, logits.grad, "d_logits")
    # mse + straight-through
    a = torch.randn(2, 3, 8, 8, generator=g).requires_grad_(True)
    b = torch.randn(2, 3, 8, 8, generator=g)
    F.mse_loss(a, b).backward()
    ac = a.detach().cuda().requires_grad_(True)
    l = ops.mse_loss(ac, b.cuda())
    (l * 1.0).backward()
    assert abs(l.item() - F.mse_loss(a, b).item()) < 1e-5
    assert_close(ac.grad, a.grad, "mse grad")
    e, q = torch.randn(2, 4, 3, 3, generator=g), torch.randn(2, 4, 3, 3, generator=g)
    ec = e.cuda().requires_grad_(True)
    st = ops.straight_through(ec, q.cuda())
    assert torch.equal(st.cpu(), e + (q - e))
    st.sum().backward()
    assert torch.equal(ec.grad.cpu(), torch.ones_like(e))


def test_ops_refuse_cpu_tensors():
    from igm_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.relu(torch.zeros(1, 2, 3, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.conv2d(torch.zeros(1, 2, 3, 3), torch.zeros(2, 2, 1, 1))
