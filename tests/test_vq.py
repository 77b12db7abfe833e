"""VQ quantiser: oracle vs reference/golden on CPU; CUDA kernel vs oracle on the GPU.
Index bar: bit-exact wherever the exact fp64 top-2 distance gap exceeds 1e-5
relative (below that the fp32 reference itself disagrees with exact arithmetic); losses/grads 1e-3."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import ref_loader, vq_oracle
from tests._util import GOLDEN, assert_close

_spec = importlib.util.spec_from_file_location("make_golden_vq", os.path.join(GOLDEN, "make_golden_vq.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)


def _golden(case):
    return dict(np.load(os.path.join(GOLDEN, f"vq_{case}.npz")))


def _oracle_all(z, emb):
    zz = z.clone().requires_grad_(True)
    e = emb.clone().requires_grad_(True)
    quant, vq_loss, commit, idx = vq_oracle.vq_forward(zz, e, 0.25)
    (vq_loss + 0.5 * commit + (quant * quant).sum() * 1e-3).backward()
    return quant.detach(), vq_loss.detach(), commit.detach(), idx, zz.grad, e.grad


@pytest.mark.parametrize("case", list(mg.CASES))
def test_oracle_matches_golden(case):
    z, emb = mg.inputs(case)
    g = _golden(case)
    quant, vq_loss, commit, idx, dz, de = _oracle_all(z, emb)
    assert np.array_equal(idx.numpy(), g["idx"])
    assert abs(vq_loss.item() - g["vq_loss"]) <= 1e-6 * abs(g["vq_loss"])
    assert abs(commit.item() - g["commit_loss"]) <= 1e-6 * abs(g["commit_loss"])
    assert_close(dz, g["dz"], "dz", 1e-6)
    assert abs(de.norm().item() - g["d_emb_norm"]) <= 1e-5 * g["d_emb_norm"]


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_oracle_bit_exact_vs_live_reference():
    ref = ref_loader.load("vqvae")
    z, emb = mg.inputs("normal")
    vq = ref.VectorQuantizer(512, 64, 0.25)
    with torch.no_grad():
        vq.embedding.copy_(emb)
    q, l1, l2 = vq(z)
    oq, ol1, ol2, _ = vq_oracle.vq_forward(z, emb, 0.25)
    assert torch.equal(q, oq) and torch.equal(l1, ol1) and torch.equal(l2, ol2)
    torch.manual_seed(3)
    a = ref.VectorQuantizer(512, 64, 0.25).embedding
    assert a.shape == (512, 64) and float(a.abs().max()) <= 1 / 512


def test_mirror_state_dict():
    import igm_b200
    torch.manual_seed(0)
    m = igm_b200.VectorQuantizer(512, 64, 0.25)
    assert list(m.state_dict().keys()) == ["embedding"] and m.embedding.shape == (512, 64)
    if ref_loader.available():
        ref = ref_loader.load("vqvae")
        torch.manual_seed(0)
        r = ref.VectorQuantizer(512, 64, 0.25)
        assert torch.equal(r.embedding, m.embedding)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 64, 2, 2))


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(mg.CASES))
def test_gpu_vq_matches_oracle(case):
    import igm_b200
    z, emb = mg.inputs(case)
    N, D, H, W, K, kind = mg.CASES[case]
    g = _golden(case)
    quant, vq_loss, commit, idx, dz, de = _oracle_all(z, emb)
    m = igm_b200.VectorQuantizer(K, D, 0.25).cuda()
    with torch.no_grad():
        m.embedding.copy_(emb.cuda())
    zc = z.cuda().requires_grad_(True)
    q, l1, l2 = m(zc)
    (l1 + 0.5 * l2 + (q * q).sum() * 1e-3).backward()
    got = m.last_indices.cpu()
    gap = vq_oracle.top2_gap(z, emb)
    decisive = gap > 1e-5
    n_bad = int((got[decisive] != idx[decisive]).sum())
    assert n_bad == 0, f"{n_bad} index mismatches among {int(decisive.sum())} decisive vectors"
    assert int((got[decisive] != torch.from_numpy(g['idx'])[decisive]).sum()) == 0
    # where indices agree the gathered codes are bit-identical
    same = (got == idx)
    qr = q.detach().cpu().reshape(N, D, -1).permute(0, 2, 1).reshape(-1, D)
    qo = quant.reshape(N, D, -1).permute(0, 2, 1).reshape(-1, D)
    assert torch.equal(qr[same], qo[same])
    assert abs(l1.item() - vq_loss.item()) <= 1e-3 * abs(vq_loss.item())
    assert abs(l2.item() - commit.item()) <= 1e-3 * abs(commit.item())
    if bool(same.all()):
        assert_close(zc.grad.cpu(), dz, "dz")
        assert_close(m.embedding.grad.cpu(), de, "d_embedding")


@pytest.mark.gpu
def test_gpu_vq_first_index_tie_break_and_full_size():
    import igm_b200
    # duplicated codes: the first copy must win (torch.argmin semantics)
    g = torch.Generator().manual_seed(1)
    emb = torch.randn(64, 64, generator=g)
    emb[40] = emb[7]
    emb[63] = emb[7]
    z = emb[7].reshape(1, 64, 1, 1).repeat(2, 1, 3, 3).contiguous()
    m = igm_b200.VectorQuantizer(64, 64, 0.25).cuda()
    with torch.no_grad():
        m.embedding.copy_(emb.cuda())
    m(z.cuda())
    assert bool((m.last_indices.cpu() == 7).all())
    # BASELINE config size (32 images x 32x32 latents, K=512): every vector gets its nearest code
    z = torch.randn(32, 64, 32, 32, generator=g)
    emb = torch.randn(512, 64, generator=g)
    m = igm_b200.VectorQuantizer(512, 64, 0.25).cuda()
    with torch.no_grad():
        m.embedding.copy_(emb.cuda())
    q, l1, l2 = m(z.cuda())
    idx = m.last_indices
    zr = z.cuda().reshape(32, 64, -1).permute(0, 2, 1).reshape(-1, 64)
    d_sel = (zr - m.embedding.detach()[idx]).norm(dim=1)
    d_min = torch.cdist(zr, m.embedding.detach()).min(dim=1).values
    assert float((d_sel - d_min).abs().max()) <= 1e-4 * float(d_min.max())
    assert abs(l2.item() - 0.25 * l1.item()) < 1e-6
