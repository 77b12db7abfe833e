"""Host-side mirror of the reference PixelCNN (src/models/pixelcnn.py:85-230), backed by the
incremental CUDA engine of csrc/pixelcnn.cu.

Same constructor, ``state_dict`` keys (masks, ``log2`` buffer included) and default initialisation as
the reference; ``forward`` (logits [N,256,C,H,W]) and ``sample`` run in ONE persistent kernel that walks
the raster once instead of re-running the network per pixel.  ``sample`` draws with the kernel's own
Philox stream (``torch.multinomial``'s stream cannot be reproduced inside a kernel); pass ``uniforms``
([H*W, N*C]) or ``greedy=True`` for reproducible / parity runs.  No CPU fallback.

Training: when autograd is recording, ``forward`` / ``calc_likelihood`` run layer by layer on the CUDA
operators of ``ops`` (masked convolutions incl. the in-place weight masking of :23, gates, ELU, the
256-way cross entropy) so ``training_step(...).backward()`` fills every ``.grad`` like the reference.
Class conditioning (``class_condition=True``) is supported on both paths.
"""
import ctypes as C
from functools import partial

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib, ops
from .ddpm import ValidationResult, _Holder, _LightningModule, _HAVE_LIGHTNING

DILATIONS = (1, 2, 1, 4, 1, 2, 1, 4, 1, 2, 1)   # reference pixelcnn.py:108-122


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _masked_conv(c_in, c_out, mask, dilation=1):
    """Parameter holder shaped like reference MaskedConvolution (:12-24): buffer `mask` + `conv`."""
    m = _Holder()
    m.register_buffer("mask", mask)
    kh, kw = mask.shape
    pad = (dilation * (kh - 1) // 2, dilation * (kw - 1) // 2)
    m.conv = nn.Conv2d(c_in, c_out, (kh, kw), padding=pad, dilation=dilation)
    return m


def _vmask(k, center):
    m = torch.ones(k, k)
    m[k // 2 + 1:, :] = 0
    if center:
        m[k // 2] = 0
    return m


def _hmask(k, center):
    m = torch.ones(1, k)
    m[0, k // 2 + 1:] = 0
    if center:
        m[0, k // 2] = 0
    return m


def _gated(channels, dilation=1, cond_channel=None):
    g = _Holder()
    g.horiz_conv = _masked_conv(channels, 2 * channels, _hmask(3, False), dilation)
    g.vert_conv = _masked_conv(channels, 2 * channels, _vmask(3, False), dilation)
    g.conv1x1_1 = nn.Conv2d(2 * channels, 2 * channels, 1)
    g.conv1x1_2 = nn.Conv2d(channels, channels, 1)
    if cond_channel is not None:   # reference :58-62, same registration order
        g.cond_proj_vert1 = nn.Conv2d(cond_channel, channels, kernel_size=1, bias=False)
        g.cond_proj_vert2 = nn.Conv2d(cond_channel, channels, kernel_size=1, bias=False)
        g.cond_proj_horiz1 = nn.Conv2d(cond_channel, channels, kernel_size=1, bias=False)
        g.cond_proj_horiz2 = nn.Conv2d(cond_channel, channels, kernel_size=1, bias=False)
    return g


def _mconv(m, x):
    """MaskedConvolution.forward (:22-24): mask the weight IN PLACE, then convolve."""
    c = m.conv
    c.weight.data *= m.mask
    return ops.conv2d(x, c.weight, c.bias, 1, c.padding, c.dilation)


class PixelCNN(_LightningModule):
    """Drop-in for reference ``PixelCNN``."""

    def __init__(self, datamodule, hidden_dim, class_condition=False, n_classes=None, lr=1e-3):
        super().__init__()
        self.width, self.height, self.channels = datamodule.width, datamodule.height, datamodule.channels
        self.input_normalize = datamodule.transforms.normalize
        self.output_act = "tanh" if self.input_normalize else "sigmoid"
        if _HAVE_LIGHTNING:
            self.save_hyperparameters(ignore=["datamodule"])
        else:
            self.save_hyperparameters(hidden_dim=hidden_dim, class_condition=class_condition, n_classes=n_classes, lr=lr)
        self.conv_vstack = _masked_conv(self.channels, hidden_dim, _vmask(5, True))
        self.conv_hstack = _masked_conv(self.channels, hidden_dim, _hmask(5, True))
        self.conv_layers = nn.ModuleList([_gated(hidden_dim, d, n_classes if class_condition else None) for d in DILATIONS])
        self.conv_out = nn.Conv2d(hidden_dim, self.channels * 256, kernel_size=1, padding=0)
        self.register_buffer("log2", torch.log(torch.tensor(2, dtype=torch.float32)))
        self._packed = None
        self._packed_key = None

    # ---- weight packing: every matrix as [K][N] with only the LIVE (unmasked) taps -------------
    def _pack(self):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed is not None and self._packed_key == key:
            return self._packed
        Hd = self.hparams.hidden_dim
        parts = []

        def mat(w):   # pad every block to a multiple of 4 floats like the C side
            w = w.reshape(-1)
            pad = (-w.numel()) % 4
            return torch.cat([w, w.new_zeros(pad)]) if pad else w

        with torch.no_grad():
            w = self.conv_vstack.conv.weight[:, :, :2, :]           # [Hd, C, 2, 5] live rows
            parts += [mat(w.permute(2, 3, 1, 0)), mat(self.conv_vstack.conv.bias)]
            w = self.conv_hstack.conv.weight[:, :, 0, :2]           # [Hd, C, 2] live cols
            parts += [mat(w.permute(2, 1, 0)), mat(self.conv_hstack.conv.bias)]
            for g in self.conv_layers:
                w = g.vert_conv.conv.weight[:, :, :2, :]            # [2Hd, Hd, 2, 3]
                parts += [mat(w.permute(2, 3, 1, 0)), mat(g.vert_conv.conv.bias)]
                parts += [mat(g.conv1x1_1.weight[:, :, 0, 0].t()), mat(g.conv1x1_1.bias)]
                w = g.horiz_conv.conv.weight[:, :, 0, :2]           # [2Hd, Hd, 2]
                parts += [mat(w.permute(2, 1, 0)), mat(g.horiz_conv.conv.bias)]
                parts += [mat(g.conv1x1_2.weight[:, :, 0, 0].t()), mat(g.conv1x1_2.bias)]
            parts += [mat(self.conv_out.weight[:, :, 0, 0].t()), mat(self.conv_out.bias)]
            flat = torch.cat([p.contiguous().float() for p in parts]).contiguous()
        lib = _lib.load()
        lib.igm_pixelcnn_weight_floats.restype = C.c_int64
        assert flat.numel() == lib.igm_pixelcnn_weight_floats(self.channels, Hd), "weight layout mismatch"
        self._packed, self._packed_key = flat, key
        return flat

    def _cond_addends(self, y, N):
        """[11][2][N][2*Hd]: cat(cond_proj_*1(y), cond_proj_*2(y)) per layer for the vertical, then the horizontal gate."""
        y4 = y.reshape(N, self.hparams.n_classes, 1, 1).float()
        rows = []
        for g in self.conv_layers:
            for a, b in ((g.cond_proj_vert1, g.cond_proj_vert2), (g.cond_proj_horiz1, g.cond_proj_horiz2)):
                rows.append(torch.cat([ops.conv2d(y4, a.weight).reshape(N, -1), ops.conv2d(y4, b.weight).reshape(N, -1)], 1))
        return rows

    def _layers(self, x, y=None):
        """Layer-by-layer forward on the CUDA operators (autograd-capable); returns conv_out [N, 256*C, H, W]."""
        N = x.shape[0]
        x = x.float()
        v = _mconv(self.conv_vstack, x)
        h = _mconv(self.conv_hstack, x)
        cond = self._cond_addends(y, N) if y is not None else None
        for i, g in enumerate(self.conv_layers):
            vc = _mconv(g.vert_conv, v)
            v = ops.gate_tanh_sigmoid(vc, None if cond is None else cond[2 * i])
            hz = _mconv(g.horiz_conv, h)
            hz = ops.conv2d(vc, g.conv1x1_1.weight, g.conv1x1_1.bias, residual=hz)       # :74
            hg = ops.gate_tanh_tanh(hz, None if cond is None else cond[2 * i + 1])      # :77 (tanh*tanh, sic)
            h = ops.conv2d(hg, g.conv1x1_2.weight, g.conv1x1_2.bias, residual=h)         # :80
        return ops.conv2d(ops.elu(h), self.conv_out.weight, self.conv_out.bias)

    def _recording(self):
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def _run(self, img, mode, uniforms=None, skip=None, want_logits=False, seed=0, cond=None):
        if img.device.type != "cuda":
            raise RuntimeError("libigm_b200 runs on CUDA (B200, sm_100a) only; there is no CPU fallback")
        lib = _lib.load()
        N, Cc, H, W = img.shape
        Hd = self.hparams.hidden_dim
        wts = self._pack()
        logits = torch.empty(N, 256, Cc, H, W, device=img.device) if want_logits else None
        n_ws = lib.igm_pixelcnn_workspace_floats(N, Cc, H, W, Hd)
        ws = torch.empty(int(n_ws), device=img.device)
        if cond is not None:
            with torch.no_grad():
                cond = torch.stack(self._cond_addends(cond.to(img.device), N)).contiguous()   # [22, N, 2*Hd]
        rc = lib.igm_pixelcnn_run(_ptr(wts), _ptr(img), _ptr(uniforms), _ptr(skip), _ptr(cond), _ptr(logits), _ptr(ws),
                                  C.c_uint64(seed), N, Cc, H, W, Hd, mode, int(bool(self.input_normalize)), _stream())
        _lib.check(None, rc)
        return logits

    def forward(self, x, y=None):
        """Logits [N, 256, C, H, W] (reference :128-154) via the teacher-forced raster walk."""
        if self._recording():
            out = self._layers(x, y)
            return out.reshape(out.shape[0], 256, out.shape[1] // 256, out.shape[2], out.shape[3])
        for m in [self.conv_vstack, self.conv_hstack] + [c for g in self.conv_layers for c in (g.horiz_conv, g.vert_conv)]:
            m.conv.weight.data *= m.mask          # the reference masks in place on every forward (:23)
        return self._run(x.contiguous().float().clone(), 2, want_logits=True, cond=y)

    def calc_likelihood(self, x, label=None):
        target = ((x + 1) / 2 * 255).to(torch.long) if self.input_normalize else (x * 255).to(torch.long)
        if self._recording():
            nll = ops.cross_entropy_256(self._layers(x, label), target)      # (N, C, H, W)
        else:
            pred = self.forward(x, label)
            out = pred.reshape(pred.shape[0], -1, pred.shape[3], pred.shape[4])
            nll = ops.cross_entropy_256(out, target)
        return (nll.mean(dim=[1, 2, 3]) / self.log2).mean()

    @torch.no_grad()
    def sample(self, img_shape, cond=None, img=None, uniforms=None, greedy=False, seed=None):
        """reference :167-195.  Pixels equal to -1 are generated, the others kept (:185)."""
        dev = self.conv_out.weight.device
        if img is None:
            img = torch.zeros(img_shape, dtype=torch.float32, device=dev) - 1
        else:
            img = img.to(dev).float().contiguous()
        skip = (img != -1).all(dim=0).all(dim=0).to(torch.uint8).contiguous()   # [H, W]: every sample filled
        if uniforms is not None:
            uniforms = uniforms.to(dev).float().contiguous()
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        self._run(img, 1 if greedy else 0, uniforms=uniforms, skip=skip, seed=seed, cond=cond)
        return img

    def configure_optimizers(self):
        optimizer = torch.optim.Adam(self.parameters(), lr=self.hparams.lr)
        scheduler = torch.optim.lr_scheduler.StepLR(optimizer, 1, gamma=0.99)
        return [optimizer], [scheduler]

    def _likelihood_of_batch(self, batch):
        img, label = batch
        if self.hparams.class_condition:
            label = F.one_hot(label, num_classes=self.hparams.n_classes).to(torch.float32)
            return self.calc_likelihood(img, label)
        return self.calc_likelihood(img)

    def training_step(self, batch, batch_idx):
        loss = self._likelihood_of_batch(batch)
        self.log("train_bpd", loss)
        return loss

    def validation_step(self, batch, batch_idx):
        img, label = batch
        N, Cc, H, W = img.shape
        loss = self._likelihood_of_batch(batch)
        self.log("val_bpd", loss)
        sample_img = None
        if batch_idx == 0:
            if self.hparams.class_condition:
                nc = self.hparams.n_classes
                sample_label = torch.arange(nc, device=img.device).reshape(nc, 1).repeat(1, 8)
                sample_label = F.one_hot(sample_label, num_classes=nc).to(torch.float32)
                sample_img = self.sample((nc * 8, Cc, H, W), cond=sample_label)
            else:
                sample_img = self.sample(img.shape)
        return ValidationResult(real_image=img, fake_image=sample_img)
