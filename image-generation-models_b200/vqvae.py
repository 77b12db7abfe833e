"""Host-side mirror of the reference vector quantiser, backed by libigm_b200.so.

``VectorQuantizer(num_embeddings, latent_dim, commitment_weight)`` keeps the constructor, the
``embedding`` parameter (same U(-1/K, 1/K) init, same state_dict key) and the
``forward(z) -> (quant_z, vq_loss, commit_loss)`` contract of reference
``src/models/vqvae.py:13-43``; the distance search, gather, both MSE losses and their gradients
run in the fused CUDA kernels of csrc/vq.cu.  No CPU fallback.
"""
import ctypes as C
import itertools

import torch
from torch import nn

from . import _lib, ops
from .ddpm import FusedAdam, ValidationResult, _HAVE_LIGHTNING, _Holder, _LightningModule  # noqa: F401


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _VQFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, embedding, beta):
        if z.device.type != "cuda":
            raise RuntimeError("libigm_b200 runs on CUDA (B200, sm_100a) only; there is no CPU fallback")
        if z.dtype != torch.float32 or embedding.dtype != torch.float32:
            raise TypeError("libigm_b200 computes in fp32")
        lib = _lib.load()
        emb = embedding.contiguous()
        N, D, H, W = z.shape
        # a channels_last z (what the CUDA encoder produces) already is the [N*H*W, D] matrix of :27-31:
        # address it as N*H*W "images" of one vector each instead of copying it to NCHW
        nhwc = z.is_contiguous(memory_format=torch.channels_last) and not z.is_contiguous()
        if nhwc:
            n_img, hw = N * H * W, 1
            quant = torch.empty_like(z, memory_format=torch.channels_last)
        else:
            z = z.contiguous()
            n_img, hw = N, H * W
            quant = torch.empty_like(z)
        K = emb.shape[0]
        idx = torch.empty(N * H * W, dtype=torch.int64, device=z.device)
        losses = torch.empty(2, dtype=torch.float32, device=z.device)
        ws = torch.empty(lib.igm_vq_workspace_floats(n_img, hw), dtype=torch.float32, device=z.device)
        rc = lib.igm_vq_forward(_ptr(z), _ptr(emb), _ptr(idx), _ptr(quant), _ptr(losses), n_img, D, hw, K, float(beta),
                                _ptr(ws), _stream())
        _lib.check(None, rc)
        ctx.save_for_backward(z, emb, idx)
        ctx.beta = float(beta)
        ctx.geom = (n_img, hw, nhwc)
        ctx.mark_non_differentiable(idx)
        return quant, losses[0], losses[1], idx

    @staticmethod
    def backward(ctx, d_quant, d_vq, d_commit, _d_idx):
        z, emb, idx = ctx.saved_tensors
        lib = _lib.load()
        N, D, H, W = z.shape
        n_img, hw, nhwc = ctx.geom
        fmt = torch.channels_last if nhwc else torch.contiguous_format
        dz = torch.empty_like(z, memory_format=fmt) if ctx.needs_input_grad[0] else None
        d_emb = torch.zeros_like(emb) if ctx.needs_input_grad[1] else None
        f = lambda t: None if t is None else t.to(torch.float32).contiguous()
        d_quant = None if d_quant is None else d_quant.to(torch.float32).contiguous(memory_format=fmt)
        d_vq, d_commit = f(d_vq), f(d_commit)
        rc = lib.igm_vq_backward(_ptr(z), _ptr(emb), _ptr(idx), _ptr(d_quant), _ptr(d_vq), _ptr(d_commit), ctx.beta,
                                 _ptr(dz), _ptr(d_emb), n_img, D, hw, emb.shape[0], _stream())
        _lib.check(None, rc)
        return dz, d_emb, None


class VectorQuantizer(nn.Module):
    """Drop-in for reference ``VectorQuantizer`` (src/models/vqvae.py:13-43)."""

    def __init__(self, num_embeddings, latent_dim, commitment_weight) -> None:
        super().__init__()
        self.embedding = nn.Parameter(
            torch.zeros(num_embeddings, latent_dim).uniform_(-1 / num_embeddings, 1 / num_embeddings))
        self.latent_dim = latent_dim
        self.commitment_weight = commitment_weight

    def forward(self, z):
        quant_z, vq_loss, commit_loss, idx = _VQFn.apply(z, self.embedding, self.commitment_weight)
        self.last_indices = idx   # [N*H*W] int64, the code chosen for every latent vector
        return quant_z, vq_loss, commit_loss


# ---------------------------------------------------------------------------
# encoder / decoder (reference src/networks/vqvae.py) on the generic CUDA conv operators
# ---------------------------------------------------------------------------
class ResidualLayer(nn.Module):
    """reference networks/vqvae.py:5-26.  The reference's in-place ReLU mutates the residual input, so the
    block computes relu(x) + f(relu(x)); that (not x + f(relu(x))) is what is reproduced here."""

    def __init__(self, in_dim, h_dim, res_h_dim):
        super().__init__()
        self.res_block = nn.Sequential(
            _Holder(), nn.Conv2d(in_dim, res_h_dim, kernel_size=3, stride=1, padding=1, bias=False),
            _Holder(), nn.Conv2d(res_h_dim, h_dim, kernel_size=1, stride=1, bias=False))

    def forward(self, x):
        h = ops.relu(x)
        f = ops.conv2d(h, self.res_block[1].weight, None, 1, 1)
        return ops.conv2d(ops.relu(f), self.res_block[3].weight, None, 1, 0, residual=h)


class ResidualStack(nn.Module):
    """reference networks/vqvae.py:29-49: ONE weight-tied layer applied n times, then ReLU."""

    def __init__(self, in_dim, h_dim, res_h_dim, n_res_layers):
        super().__init__()
        self.n_res_layers = n_res_layers
        self.stack = nn.ModuleList([ResidualLayer(in_dim, h_dim, res_h_dim)] * n_res_layers)

    def forward(self, x):
        for layer in self.stack:
            x = layer(x)
        return ops.relu(x)


class Encoder(nn.Module):
    """reference networks/vqvae.py:52-96."""

    def __init__(self, input_channel, output_channel, n_res_layers=3, res_h_dim=128):
        super().__init__()
        self.conv_stack = nn.Sequential(
            nn.Conv2d(input_channel, output_channel // 2, kernel_size=4, stride=2, padding=1), _Holder(),
            nn.Conv2d(output_channel // 2, output_channel, kernel_size=4, stride=2, padding=1), _Holder(),
            nn.Conv2d(output_channel, output_channel, kernel_size=3, stride=1, padding=1),
            ResidualStack(output_channel, output_channel, res_h_dim, n_res_layers))

    def forward(self, x):
        cs = self.conv_stack
        x = ops.relu(ops.conv2d(x, cs[0].weight, cs[0].bias, 2, 1))
        x = ops.relu(ops.conv2d(x, cs[2].weight, cs[2].bias, 2, 1))
        x = ops.conv2d(x, cs[4].weight, cs[4].bias, 1, 1)
        return cs[5](x)


class Decoder(nn.Module):
    """reference networks/vqvae.py:99-136."""

    def __init__(self, input_channel, output_channel, h_dim=128, n_res_layers=3, res_h_dim=128):
        super().__init__()
        self.inverse_conv_stack = nn.Sequential(
            nn.ConvTranspose2d(input_channel, h_dim, kernel_size=3, stride=1, padding=1),
            ResidualStack(h_dim, h_dim, res_h_dim, n_res_layers),
            nn.ConvTranspose2d(h_dim, h_dim // 2, kernel_size=4, stride=2, padding=1), _Holder(),
            nn.ConvTranspose2d(h_dim // 2, output_channel, kernel_size=4, stride=2, padding=1))

    def forward(self, x):
        cs = self.inverse_conv_stack
        x = ops.conv_transpose2d(x, cs[0].weight, cs[0].bias, 1, 1)
        x = cs[1](x)
        x = ops.relu(ops.conv_transpose2d(x, cs[2].weight, cs[2].bias, 2, 1))
        return ops.conv_transpose2d(x, cs[4].weight, cs[4].bias, 2, 1)


class GradArena:
    """Flat fp32 gradient arena over a set of parameters: every ``p.grad`` is a view into ONE buffer, so the
    data-parallel exchange is a single all-reduce instead of one per tensor (same scheme as the DDPM mirror's
    ``Unet._flat_grad``).  autograd accumulates in place into an existing ``.grad``, so the views survive backward."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4
        p0 = self.params[0]
        self.flat = torch.zeros(off, dtype=torch.float32, device=p0.device)
        self.attach()

    def attach(self):
        """Re-home every ``.grad`` in the arena (a grad that was dropped or replaced is copied in first)."""
        for p, off in zip(self.params, self.offsets):
            view = self.flat[off:off + p.numel()].view(p.shape)
            g = p.grad
            if g is None:
                view.zero_()
            elif g.data_ptr() != view.data_ptr():
                view.copy_(g)
            else:
                continue
            p.grad = view

    def zero(self):
        self.flat.zero_()
        self.attach()

    def all_reduce_mean(self):
        import torch.distributed as dist
        self.attach()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.mul_(1.0 / dist.get_world_size())


class ArenaAdam(torch.optim.Adam):
    """``torch.optim.Adam`` whose ``zero_grad`` keeps the gradients attached to a ``GradArena`` (one memset)."""

    def __init__(self, arena: GradArena, **kw):
        super().__init__(arena.params, **kw)
        self.arena = arena

    def zero_grad(self, set_to_none: bool = False):
        self.arena.zero()


def _dist_world():
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _build(cfg, default_cls, **kw):
    """Stand-in for hydra.utils.instantiate(cfg, **kw) restricted to the two network classes."""
    if isinstance(cfg, nn.Module):
        return cfg
    extra = {}
    if cfg:
        extra = {k: v for k, v in dict(cfg).items() if k not in ("_target_", "input_channel", "output_channel") and v is not None}
    return default_cls(**kw, **extra)


class VQVAE(_LightningModule):
    """Drop-in for reference ``VQVAE`` (src/models/vqvae.py:46-138)."""

    def __init__(self, datamodule, encoder=None, decoder=None, latent_dim=100, lr: float = 0.0002, b1: float = 0.5,
                 b2: float = 0.999, num_embeddings: int = 512, beta: float = 0.25, optim="adam", **kwargs):
        super().__init__()
        self.width, self.height, self.channels = datamodule.width, datamodule.height, datamodule.channels
        self.input_normalize = datamodule.transforms.normalize
        self.output_act = "tanh" if self.input_normalize else "sigmoid"
        if _HAVE_LIGHTNING:
            self.save_hyperparameters(ignore=["datamodule"])
        else:
            self.save_hyperparameters(latent_dim=latent_dim, lr=lr, b1=b1, b2=b2, num_embeddings=num_embeddings,
                                      beta=beta, optim=optim, **kwargs)
        # same construction order as the reference (:64-70): decoder, encoder, quantiser
        self.decoder = _build(decoder, Decoder, input_channel=latent_dim, output_channel=self.channels)
        self.encoder = _build(encoder, Encoder, input_channel=self.channels, output_channel=latent_dim)
        self.vector_quntizer = VectorQuantizer(num_embeddings, latent_dim, beta)   # sic
        self.latent_w = self.width // 4
        self.latent_h = self.height // 4
        self.latent_size = self.latent_h * self.latent_w

    def forward(self, imgs):
        z = self.encoder(imgs)
        quant_z, _, _ = self.vector_quntizer(z)
        out = self.decoder(quant_z)
        return out.reshape(out.shape[0], self.channels, self.height, self.width)

    def training_step(self, batch, batch_idx):
        imgs, _ = batch
        encoder_z = self.encoder(imgs)
        quant_z, vq_loss, commit_loss = self.vector_quntizer(encoder_z)
        decoder_z = ops.straight_through(encoder_z, quant_z)            # encoder_z + (quant_z - encoder_z).detach() (:103)
        fake_imgs = self.decoder(decoder_z).reshape(-1, self.channels, self.height, self.width)
        recon_loss = ops.mse_loss(fake_imgs, imgs)
        total_loss = recon_loss + vq_loss + self.hparams.beta * commit_loss   # beta applied twice, like the reference (:39, :110)
        self.log("train_loss/vq_loss", vq_loss)
        self.log("train_loss/recon_loss", recon_loss)
        self.log("train_loss/commit_loss", commit_loss)
        return total_loss

    def configure_optimizers(self):
        params = list(itertools.chain(self.encoder.parameters(), self.decoder.parameters(),
                                      self.vector_quntizer.parameters()))
        self._arena = GradArena(params)
        return ArenaAdam(self._arena, lr=self.hparams.lr, betas=(self.hparams.b1, self.hparams.b2))

    ddp_sync = True   # set False when the module is wrapped in DistributedDataParallel (which reduces by itself)

    def on_after_backward(self):
        """Lightning hook (called right after ``loss.backward()``; plain loops call it themselves): the data-parallel
        exchange, ONE all-reduce (mean) over the flat gradient arena -- encoder, decoder and codebook gradients together
        (2.2 MB at BASELINE.json configs[4]).  Replicas must start from identical parameters
        (``sync_parameters``)."""
        if _dist_world() > 1 and self.ddp_sync:
            if getattr(self, "_arena", None) is None:
                self._arena = GradArena(list(self.parameters()))
            self._arena.all_reduce_mean()

    def sync_parameters(self):
        """Broadcast rank 0's parameters to every replica (one flat buffer)."""
        if _dist_world() > 1:
            import torch.distributed as dist
            ps = list(self.parameters())
            flat = torch.cat([p.detach().reshape(-1) for p in ps])
            dist.broadcast(flat, src=0)
            off = 0
            with torch.no_grad():
                for p in ps:
                    p.copy_(flat[off:off + p.numel()].view_as(p))
                    off += p.numel()

    def validation_step(self, batch, batch_idx):
        imgs, labels = batch
        recon_imgs = self.forward(imgs)
        self.log("val/recon_loss", torch.nn.functional.mse_loss(imgs, recon_imgs))
        return ValidationResult(real_image=imgs, recon_image=recon_imgs)
