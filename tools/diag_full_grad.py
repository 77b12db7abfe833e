"""Diagnostic of the full-size gradient property (tests/test_zz_gpu_full_size.py): who is right, and by how much?

For one config at its full per-GPU batch (L2 loss) this prints the error of
  g_full      gradient of the batch-mean loss, one backward at B
  g_quarters  mean of the four quarter-batch gradients (accumulated in .grad)
  g_eager32   the oracle functions on the same GPU in fp32 (cuDNN / cuBLAS, TF32 off) -- the "PyTorch eager" path
against the oracle run in fp64 on the same GPU (the arbiter), plus g_full vs g_quarters, the ten worst parameters,
and B < max_batch runs (B = 32, 80 inside an engine planned for 128) against a freshly planned engine.

    python tools/diag_full_grad.py [cifar10_b128|celeba64_b32]       (env switches: IGM_WGRAD_STREAM=0, IGM_PDL=0, ...)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import igm_b200  # noqa: E402
from oracle import ddpm_oracle as O  # noqa: E402
from tests.test_zz_gpu_full_size import CONFIGS, T, _build, _inputs  # noqa: E402


def rel(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item(), ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


def oracle_grads(params, spec, x, t, noise, dtype):
    dev = torch.device("cuda")
    p = {k: v.to(dev, dtype).requires_grad_(True) for k, v in params.items()}
    buf = {k: v.to(dev) for k, v in O.diffusion_buffers(T).items()}
    loss = O.p_losses(p, spec, buf, x.to(dev, dtype), t.to(dev), noise.to(dev, dtype), "l2")
    g = torch.autograd.grad(loss, list(p.values()))
    return loss.item(), g


def flat_like(unet, grads):
    out = torch.zeros_like(unet._flat_grad, dtype=grads[0].dtype)
    for (name, off, shape), g in zip(unet._layout, grads):
        out[off:off + g.numel()] = g.reshape(-1)
    return out


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cifar10_b128"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    env = {k: v for k, v in os.environ.items() if k.startswith("IGM_")}
    print(f"== {cfg} env={env}")
    spec, params, unet, gd = _build(cfg, "l2")
    x, t, noise = _inputs(cfg, seed=99)
    B = x.shape[0]
    xc, tc, nc = x.cuda(), t.cuda(), noise.cuda()

    def ours(sl):
        unet._flat_grad.zero_()
        loss = gd.p_losses(xc[sl].contiguous(), tc[sl].contiguous(), nc[sl].contiguous())
        loss.backward()
        torch.cuda.synchronize()
        return loss.item(), unet._flat_grad.clone()

    l_full, g_full = ours(slice(0, B))
    l_full2, g_full2 = ours(slice(0, B))
    print("repeat of the full backward: rel-L2 %.3e max-rel %.3e" % rel(g_full2, g_full))
    q = B // 4
    acc = torch.zeros_like(g_full)
    lq = []
    for i in range(4):
        l, g = ours(slice(i * q, (i + 1) * q))
        acc += g
        lq.append(l)
    g_quarters = acc / 4

    l64, g64 = oracle_grads(params, spec, x, t, noise, torch.float64)
    ref = flat_like(unet, g64)
    l32, g32 = oracle_grads(params, spec, x, t, noise, torch.float32)
    eager = flat_like(unet, g32)
    print(f"loss: ours {l_full:.8f}  quarters {sum(lq) / 4:.8f}  fp64 {l64:.8f}  eager32 {l32:.8f}")
    print("g_full     vs fp64    : rel-L2 %.3e max-rel %.3e" % rel(g_full, ref))
    print("g_quarters vs fp64    : rel-L2 %.3e max-rel %.3e" % rel(g_quarters, ref))
    print("g_eager32  vs fp64    : rel-L2 %.3e max-rel %.3e" % rel(eager, ref))
    print("g_quarters vs g_full  : rel-L2 %.3e max-rel %.3e" % rel(g_quarters, g_full))
    rows = []
    for name, off, shape in unet._layout:
        n = 1
        for s in shape:
            n *= s
        a, b, r = g_full[off:off + n], g_quarters[off:off + n], ref[off:off + n]
        rows.append((rel(b, a)[0], rel(a, r)[0], rel(b, r)[0], name, n))
    rows.sort(reverse=True)
    print("worst parameters (quarters-vs-full, full-vs-fp64, quarters-vs-fp64):")
    for r in rows[:10]:
        print("   %.3e  %.3e  %.3e  %s [%d]" % r)
    worst64 = max(rows, key=lambda r: r[1])
    print("worst full-vs-fp64 parameter: %.3e %s" % (worst64[1], worst64[3]))

    # B < max_batch inside the engine planned for B (the epoch-tail batch) vs a freshly planned engine
    for b in (32, 80):
        if b >= B:
            continue
        sl = slice(0, b)
        _, g_in_big = ours(sl)
        spec2, params2, unet2, gd2 = _build(cfg, "l2")
        unet2._flat_grad.zero_()
        loss2 = gd2.p_losses(xc[sl].contiguous(), tc[sl].contiguous(), nc[sl].contiguous())
        loss2.backward()
        torch.cuda.synchronize()
        g_fresh = unet2._flat_grad.clone()
        l64b, g64b = oracle_grads(params, spec, x[sl], t[sl], noise[sl], torch.float64)
        refb = flat_like(unet, g64b)
        print(f"B={b} in max_batch={B} engine vs fresh engine : rel-L2 %.3e max-rel %.3e" % rel(g_in_big, g_fresh))
        print(f"B={b} in max_batch={B} engine vs fp64         : rel-L2 %.3e max-rel %.3e" % rel(g_in_big, refb))
        del unet2, gd2


if __name__ == "__main__":
    main()
