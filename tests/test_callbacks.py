"""CPU: the consumer side of the hot path -- SampleImagesCallback / get_grid_images (reference
src/callbacks/visualization.py:13-38, :141-148) and checkpoint interchange of the DDPM state_dict with the reference."""
import io
import os
from types import SimpleNamespace

import pytest
import torch

import igm_b200
from oracle import ddpm_oracle as O
from oracle import ref_loader


class _Exp:
    def __init__(self):
        self.images = {}

    def add_image(self, tag, img, global_step=None):
        self.images[tag] = (img, global_step)


@pytest.mark.parametrize("normalize", [True, False])
def test_sample_images_callback_writes_grid(tmp_path, monkeypatch, normalize):
    monkeypatch.chdir(tmp_path)
    g = torch.Generator().manual_seed(0)
    real = torch.rand(70, 3, 8, 8, generator=g) * 2 - 1 if normalize else torch.rand(70, 3, 8, 8, generator=g)
    fake = torch.rand(64, 3, 8, 8, generator=g) * 2 - 1 if normalize else torch.rand(64, 3, 8, 8, generator=g)
    out = igm_b200.ValidationResult(real_image=real, fake_image=fake, others={"diffusion": real[:16], "none": None})
    model = SimpleNamespace(input_normalize=normalize)
    exp = _Exp()
    trainer = SimpleNamespace(current_epoch=3, logger=SimpleNamespace(experiment=exp))
    cb = igm_b200.SampleImagesCallback(every_n_epochs=1)
    cb.on_validation_batch_end(trainer, model, out, None, 1)          # only batch 0 is visualised
    assert not exp.images
    cb.on_validation_batch_end(trainer, model, out, None, 0)
    assert set(exp.images) == {"images/real", "images/sample", "images/diffusion"}
    grid, step = exp.images["images/sample"]
    assert step == 3 and grid.shape == (3, 8 * 10 + 2, 8 * 10 + 2)    # 8x8 tiles of 8 px + 2 px padding
    assert float(grid.min()) >= 0.0 and float(grid.max()) <= 1.0
    assert grid[:, 0, 0].tolist() == [1.0, 1.0, 1.0]                   # pad_value = 1
    # first tile = first image mapped from [-1, 1] to [0, 1] when the datamodule normalises
    want = (fake[0] + 1) / 2 if normalize else fake[0]
    assert torch.allclose(grid[:, 2:10, 2:10], want.clamp(0, 1), atol=1e-6)
    assert exp.images["images/real"][0].shape[1] == 8 * 10 + 2         # 64 of the 70 images
    assert os.path.exists(tmp_path / "results" / "3.jpg")


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_checkpoint_interchange_with_reference():
    """A reference DDPM checkpoint loads into the mirror and the mirror's checkpoint loads into the reference
    (same 368 keys incl. the aliased diffusion_model.denoise_fn.* copies and the 12 schedule buffers)."""
    ref = ref_loader.load("ddpm")
    dm = ref_loader.datamodule_cfg(3, 32, 32)
    torch.manual_seed(0)
    rm = ref.DDPM(dm, hidden_dim=64, dim_mults=[1, 2, 4], timesteps=1000, loss_type="l1")
    mine = igm_b200.DDPM(dm, hidden_dim=64, dim_mults=(1, 2, 4), timesteps=1000, loss_type="l1")
    sd = rm.state_dict()
    assert list(sd.keys()) == list(mine.state_dict().keys()) and len(sd) == 368
    buf = io.BytesIO()
    torch.save({"state_dict": sd}, buf)            # Lightning's .ckpt is a torch.save'd dict with this key
    buf.seek(0)
    mine.load_state_dict(torch.load(buf)["state_dict"])
    for (k, a), (_, b) in zip(mine.state_dict().items(), sd.items()):
        assert a.shape == b.shape and torch.equal(a.cpu(), b), k
    # and back: perturb, save from the mirror, load into the reference
    with torch.no_grad():
        for p in mine.denoising_model.parameters():
            p.add_(0.125)
    buf = io.BytesIO()
    torch.save({"state_dict": mine.state_dict()}, buf)
    buf.seek(0)
    rm.load_state_dict(torch.load(buf)["state_dict"])
    for p, q in zip(rm.denoising_model.parameters(), mine.denoising_model.parameters()):
        assert torch.equal(p, q.cpu())
    assert torch.equal(rm.diffusion_model.betas, mine.diffusion_model.betas.cpu())


@pytest.mark.gpu
def test_validation_step_output_feeds_the_callback_from_the_device(tmp_path, monkeypatch):
    """The step after the hot path, on hardware: DDPM.validation_step (reference ddpm.py:514-521 -- q_sample of the batch at
    t = T-1 and a 64-image reverse chain on batch 0) hands DEVICE tensors to SampleImagesCallback, which must produce the
    same grids as from host copies, write results/<epoch>.jpg, and the checkpoint written afterwards must reload."""
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(0)
    T = 20
    d = igm_b200.DDPM(ref_loader.datamodule_cfg(3, 16, 16), hidden_dim=32, dim_mults=(1, 2), timesteps=T, lr=1e-4, b1=0.9,
                      b2=0.999).cuda()
    imgs = (torch.randn(8, 3, 16, 16) * 0.5).clamp(-1, 1).cuda()
    out = d.validation_step((imgs, None), 0)
    assert out.fake_image.is_cuda and out.fake_image.shape == (64, 3, 16, 16) and torch.isfinite(out.fake_image).all()
    assert out.others["diffusion"].is_cuda
    # q_sample at t = T-1 against the oracle's schedule
    buf = O.diffusion_buffers(T)
    exp = _Exp()
    trainer = SimpleNamespace(current_epoch=2, logger=SimpleNamespace(experiment=exp))
    cb = igm_b200.SampleImagesCallback()
    cb.on_validation_batch_end(trainer, d, out, (imgs, None), 0)
    assert set(exp.images) == {"images/real", "images/sample", "images/diffusion"}
    grid = exp.images["images/sample"][0]
    host_grid = igm_b200.get_grid_images(out.fake_image.cpu(), d)
    assert grid.is_cuda and torch.allclose(grid.cpu(), host_grid, atol=1e-6)
    assert os.path.exists(tmp_path / "results" / "2.jpg")
    # checkpoint round trip after validation (what ModelCheckpoint does next)
    b = io.BytesIO()
    torch.save({"state_dict": d.state_dict()}, b)
    b.seek(0)
    d2 = igm_b200.DDPM(ref_loader.datamodule_cfg(3, 16, 16), hidden_dim=32, dim_mults=(1, 2), timesteps=T).cuda()
    d2.load_state_dict(torch.load(b)["state_dict"])
    x_t = d2.diffusion_model.q_sample(imgs, torch.full((8,), T - 1, device="cuda", dtype=torch.long), noise=torch.zeros_like(imgs))
    want = buf["sqrt_alphas_cumprod"][T - 1] * imgs.cpu()
    assert torch.allclose(x_t.cpu(), want, atol=1e-6)
    with torch.no_grad():
        a = d.denoising_model(imgs, torch.full((8,), 3, device="cuda", dtype=torch.long))
        c = d2.denoising_model(imgs, torch.full((8,), 3, device="cuda", dtype=torch.long))
    assert torch.equal(a, c)
