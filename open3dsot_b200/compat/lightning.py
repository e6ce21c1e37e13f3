"""Minimal `LightningModule` surface used by the reference models (models/base_model.py:17-36,
models/bat.py:18-20,146-164): `save_hyperparameters`, `log`, `device`, `global_step`, `logger`,
`load_from_checkpoint`.  Falls back to this shim only when pytorch_lightning is not installed."""
import inspect

import torch
import torch.nn as nn

try:
    from pytorch_lightning import LightningModule  # type: ignore # noqa: F401
    HAVE_LIGHTNING = True
except Exception:  # pragma: no cover - exercised in this image
    HAVE_LIGHTNING = False

    class _NullExperiment:
        def add_scalars(self, *a, **k):
            pass

        def add_scalar(self, *a, **k):
            pass

    class _NullLogger:
        experiment = _NullExperiment()

    class LightningModule(nn.Module):
        def __init__(self):
            super().__init__()
            self.hparams = {}
            self.global_step = 0
            self.logger = _NullLogger()
            self.logged = {}

        def save_hyperparameters(self, *args, **kwargs):
            frame = inspect.currentframe().f_back
            names = inspect.getargvalues(frame)
            hp = {k: names.locals[k] for k in names.args if k != "self"}
            if names.keywords and names.keywords in names.locals:
                hp.update(names.locals[names.keywords])
            self.hparams = hp

        def log(self, name, value, **kwargs):
            self.logged[name] = value

        @property
        def device(self):
            for p in self.parameters():
                return p.device
            return torch.device("cpu")

        @classmethod
        def load_from_checkpoint(cls, checkpoint_path, map_location=None, **kwargs):
            from ..checkpoint import load_lightning_checkpoint
            ckpt = load_lightning_checkpoint(checkpoint_path, map_location=map_location)
            hp = dict(ckpt.get("hyper_parameters", {}))
            hp.update(kwargs)
            model = cls(**hp)
            model.load_state_dict(ckpt["state_dict"])
            return model
