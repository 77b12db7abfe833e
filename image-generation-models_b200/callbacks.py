"""The step after the hot path: what consumes ``validation_step``'s ``fake_image``.

Mirrors reference ``src/callbacks/visualization.py`` ``SampleImagesCallback`` (:13-38) and ``get_grid_images``
(:141-148) without the matplotlib / pytorch_lightning imports the reference module needs at import time: same
constructor, same hook name and arguments, same grid (torchvision ``make_grid``: 8 per row, white padding,
[-1, 1] -> [0, 1] when the datamodule normalises), same outputs (``images/{real,recon,sample,<key>}`` on the logger,
``results/{epoch}.jpg`` on disk).
"""
from pathlib import Path

import torch
import torchvision


def get_grid_images(imgs: torch.Tensor, model, nimgs: int = 64, nrow: int = 8) -> torch.Tensor:
    """reference visualization.py:141-148"""
    if model.input_normalize:
        return torchvision.utils.make_grid(imgs[:nimgs], nrow=nrow, normalize=True, value_range=(-1, 1), pad_value=1)
    return torchvision.utils.make_grid(imgs[:nimgs], normalize=False, nrow=nrow, pad_value=1)


class SampleImagesCallback:
    """reference visualization.py:13-38 (a pytorch_lightning.Callback when Lightning is importable)."""

    def __init__(self, batch_size=64, every_n_epochs=1):
        self.batch_size = batch_size
        self.every_n_epochs = every_n_epochs

    def on_validation_batch_end(self, trainer, pl_module, outputs, batch, batch_idx, *unused):
        if trainer.current_epoch % self.every_n_epochs != 0 or batch_idx != 0:
            return
        result_path = Path("results")
        result_path.mkdir(parents=True, exist_ok=True)
        experiment = getattr(getattr(trainer, "logger", None), "experiment", None)

        def log(tag, grid):
            if experiment is not None:
                experiment.add_image(tag, grid, global_step=trainer.current_epoch)

        log("images/real", get_grid_images(outputs.real_image, pl_module))
        if outputs.recon_image is not None:
            log("images/recon", get_grid_images(outputs.recon_image, pl_module))
        if outputs.fake_image is not None:
            fake_grid = get_grid_images(outputs.fake_image, pl_module)
            log("images/sample", fake_grid)
            torchvision.utils.save_image(fake_grid, result_path / f"{trainer.current_epoch}.jpg")
        for key, val in (outputs.others or {}).items():
            if val is not None:
                log(f"images/{key}", get_grid_images(val, pl_module))


try:  # become a real Lightning callback when Lightning is installed (it is not in this image)
    import pytorch_lightning as _pl  # type: ignore

    SampleImagesCallback = type("SampleImagesCallback", (SampleImagesCallback, _pl.Callback), {})
except Exception:  # pragma: no cover
    pass
