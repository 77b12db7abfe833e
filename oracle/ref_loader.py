"""Import the UNMODIFIED reference modules from /root/reference on CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference needs
pytorch_lightning / omegaconf / hydra / rich, none of which exist in this
image, so a handful of empty stand-in modules are registered in sys.modules
before ``src.models.*`` is imported.  No reference
source is copied: the files are executed from where they lie.  This loader is
only usable where /root/reference exists (the build container); the GPU box
relies on the committed fixtures under tests/golden/ instead.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("IGM_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "models", "ddpm.py"))


class _AttrDict(dict):
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


def _install_stubs():
    import torch
    from torch import nn

    if "pytorch_lightning" in sys.modules and getattr(sys.modules["pytorch_lightning"], "_igm_stub", False):
        return

    pl = types.ModuleType("pytorch_lightning")
    pl._igm_stub = True

    class LightningModule(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self.hparams = _AttrDict()
            self._logged = {}

        def save_hyperparameters(self, *a, **k):
            import inspect

            frame = inspect.currentframe().f_back
            args = inspect.getargvalues(frame)
            for name in args.args:
                if name not in ("self", "datamodule"):
                    self.hparams[name] = args.locals[name]
            if args.keywords and args.locals.get(args.keywords):
                self.hparams.update(args.locals[args.keywords])

        def log(self, name, value, *a, **k):
            self._logged[name] = value

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

    class _Any:
        def __init__(self, *a, **k):
            pass

    pl.LightningModule = LightningModule
    pl.LightningDataModule = _Any
    pl.Callback = _Any
    pl.Trainer = _Any
    pl.seed_everything = lambda s, **k: torch.manual_seed(s)

    util = types.ModuleType("pytorch_lightning.utilities")
    util.rank_zero_only = lambda f: f
    loggers = types.ModuleType("pytorch_lightning.loggers")
    loggers.Logger = _Any
    loggers.LightningLoggerBase = _Any
    cbs = types.ModuleType("pytorch_lightning.callbacks")
    cbs.Callback = _Any
    pl.utilities = util
    pl.loggers = loggers
    pl.callbacks = cbs

    oc = types.ModuleType("omegaconf")
    oc.DictConfig = dict
    oc.OmegaConf = type("OmegaConf", (), {"to_container": staticmethod(lambda c, **k: c)})

    hy = types.ModuleType("hydra")
    hyu = types.ModuleType("hydra.utils")

    def _instantiate(cfg, *a, **k):
        # minimal stand-in: import cfg["_target_"] and call it with the remaining keys overridden by **k
        import importlib

        kw = {key: v for key, v in dict(cfg).items() if key != "_target_"}
        kw.update(k)
        mod, _, attr = cfg["_target_"].rpartition(".")
        return getattr(importlib.import_module(mod), attr)(*a, **kw)

    hyu.instantiate = _instantiate
    hy.utils = hyu

    rich = types.ModuleType("rich")
    rich_syntax = types.ModuleType("rich.syntax")
    rich_tree = types.ModuleType("rich.tree")
    rich.syntax = rich_syntax
    rich.tree = rich_tree

    for name, mod in {
        "pytorch_lightning": pl,
        "pytorch_lightning.utilities": util,
        "pytorch_lightning.loggers": loggers,
        "pytorch_lightning.callbacks": cbs,
        "omegaconf": oc,
        "hydra": hy,
        "hydra.utils": hyu,
    }.items():
        sys.modules.setdefault(name, mod)
    try:
        import rich as _real_rich  # noqa: F401  (present in this image)
        import rich.syntax, rich.tree  # noqa: F401,E401
    except Exception:
        sys.modules.setdefault("rich", rich)
        sys.modules.setdefault("rich.syntax", rich_syntax)
        sys.modules.setdefault("rich.tree", rich_tree)


def load(module: str):
    """Return the reference module ``src.models.<module>`` executed unmodified."""
    if not available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    return importlib.import_module(f"src.models.{module}")


def datamodule_cfg(channels, height, width, normalize=True):
    """The DictConfig slice BaseModel reads (reference src/models/base.py:20-23)."""
    return _AttrDict(width=width, height=height, channels=channels,
                     transforms=_AttrDict(normalize=normalize, convert=True))
