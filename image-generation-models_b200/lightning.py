"""Lightning glue for multi-GPU runs of the mirrored modules (only meaningful where pytorch_lightning is installed).

Why a strategy is needed: Lightning's ``strategy="ddp"`` wraps the LightningModule in ``DistributedDataParallel``.
The DDPM mirror writes its parameter gradients straight into a flat arena from CUDA kernels (outside autograd), so
DDP's reducer hooks never fire and its second iteration raises "Expected to have finished reduction in the prior
iteration"; the mirror also issues its own gradient all-reduce (ddpm.py ``_allreduce``).  ``IGMDDPStrategy`` keeps
everything else of ``DDPStrategy`` (process launch, rendezvous, distributed sampler, rank-zero logging) but leaves the
module UNWRAPPED; replicas are synchronised by ``Unet.sync_parameters()`` (one broadcast of the flat arena from rank 0
when the engine is created) and by the mirror's own all-reduce of each backward's gradients.

    trainer = pl.Trainer(accelerator="gpu", devices=8, strategy=igm_b200.lightning.IGMDDPStrategy())

NOT exercised in this repository's tests: pytorch_lightning is absent from the build image.
The single-process path (``devices=1``) needs no strategy.  Without Lightning, importing this module still works and
``IGMDDPStrategy`` raises at construction.
"""
try:  # pragma: no cover - Lightning is absent from this image
    from pytorch_lightning.strategies import DDPStrategy as _DDPStrategy
    _HAVE = True
except Exception:
    _DDPStrategy = object
    _HAVE = False


class IGMDDPStrategy(_DDPStrategy):
    """``DDPStrategy`` that does not wrap the module in DistributedDataParallel."""

    strategy_name = "igm_ddp"

    def __init__(self, *args, **kwargs):
        if not _HAVE:
            raise RuntimeError("IGMDDPStrategy needs pytorch_lightning (not installed in this environment)")
        super().__init__(*args, **kwargs)

    # Lightning 2.x: DDPStrategy.configure_ddp() == self.model = self._setup_model(self.model); self._register_ddp_hooks()
    def _setup_model(self, model):  # pragma: no cover
        return model

    def _register_ddp_hooks(self):  # pragma: no cover
        return None

    def configure_ddp(self):  # pragma: no cover
        return None
