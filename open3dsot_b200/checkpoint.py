"""Reading the reference's Lightning checkpoints (`pretrained_models/*.ckpt`, written by pytorch-lightning 1.3.8 in the
legacy non-zip torch format) without pytorch_lightning / easydict installed.

Such a file pickles, besides tensors, an `easydict.EasyDict` (hyper_parameters.config) and a
`pytorch_lightning.callbacks.model_checkpoint.ModelCheckpoint` class used as a dict key.
`load_lightning_checkpoint` unpickles with a restricted `find_class`: torch / collections / numpy names resolve normally,
`easydict.EasyDict` maps to our stand-in, every other foreign global becomes an inert placeholder class.  The result has
the reference's keys: `state_dict`, `hyper_parameters`, `epoch`, `global_step`, `optimizer_states`, …
"""
import pickle

import torch

from .compat.easydict import EasyDict

_SAFE_PREFIXES = ("torch", "collections", "numpy", "builtins", "_codecs", "copyreg")


def _placeholder(module, name):
    return type(name, (), {"__module__": module, "__init__": lambda self, *a, **k: None,
                           "__setstate__": lambda self, state: None, "__reduce__": lambda self: (str, ("",))})


class _RestrictedUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "easydict" and name == "EasyDict":
            return EasyDict
        if module.split(".")[0] in _SAFE_PREFIXES:
            return super().find_class(module, name)
        return _placeholder(module, name)


class _PickleModule:
    """The `pickle_module` torch.load expects: module-like object exposing Unpickler / load / loads."""
    __name__ = "open3dsot_b200_restricted_pickle"
    Unpickler = _RestrictedUnpickler
    Pickler = pickle.Pickler
    PickleError = pickle.PickleError
    UnpicklingError = pickle.UnpicklingError

    @staticmethod
    def load(f, **kw):
        return _RestrictedUnpickler(f, **kw).load()

    @staticmethod
    def loads(b, **kw):
        import io
        return _RestrictedUnpickler(io.BytesIO(b), **kw).load()


def load_lightning_checkpoint(path, map_location="cpu"):
    return torch.load(path, map_location=map_location, pickle_module=_PickleModule, weights_only=False)


def load_reference_weights(model, path, strict=True, map_location="cpu"):
    """Load `state_dict` of a reference checkpoint into one of our models (same parameter names). Returns the ckpt dict."""
    ckpt = load_lightning_checkpoint(path, map_location=map_location)
    sd = {k: v for k, v in ckpt["state_dict"].items()}
    model.load_state_dict(sd, strict=strict)
    return ckpt
