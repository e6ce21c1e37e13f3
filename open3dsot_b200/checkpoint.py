"""Reading the reference's Lightning checkpoints (`pretrained_models/*.ckpt`, written by pytorch-lightning 1.3.8 in the
legacy non-zip torch format) without pytorch_lightning / easydict installed.

Such a file pickles, besides tensors, an `easydict.EasyDict` (hyper_parameters.config) and a
`pytorch_lightning.callbacks.model_checkpoint.ModelCheckpoint` class used as a dict key.
`load_lightning_checkpoint` unpickles with a restricted `find_class`: only an explicit whitelist of (module, name)
pairs resolves to the real object — the tensor / storage rebuild helpers, `collections.OrderedDict`, the numpy array /
scalar reconstruction helpers and a few inert builtins containers; `easydict.EasyDict` maps to our stand-in; EVERY other
global (including `builtins.eval/exec/getattr/__import__`, `os.*`, `torch.hub.*`, `numpy.testing.*`) becomes an inert
placeholder class whose construction and `__setstate__` do nothing, so a crafted file cannot call into them.  The result has
the reference's keys: `state_dict`, `hyper_parameters`, `epoch`, `global_step`, `optimizer_states`, …
"""
import pickle

import torch

from .compat.easydict import EasyDict

_STORAGES = {n for n in ("DoubleStorage", "FloatStorage", "HalfStorage", "BFloat16Storage", "LongStorage", "IntStorage",
                        "ShortStorage", "CharStorage", "ByteStorage", "BoolStorage", "UntypedStorage")}
# exact (module, name) pairs that resolve to the real object; nothing else does
_SAFE = {
    ("collections", "OrderedDict"), ("collections", "defaultdict"),
    ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_tensor"), ("torch._utils", "_rebuild_parameter"),
    ("torch", "Size"), ("torch", "device"), ("torch", "dtype"),
    ("torch.serialization", "_get_layout"),
    ("numpy.core.multiarray", "_reconstruct"), ("numpy.core.multiarray", "scalar"),
    ("numpy._core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "scalar"),
    ("numpy", "ndarray"), ("numpy", "dtype"),
    ("_codecs", "encode"),
    ("builtins", "set"), ("builtins", "frozenset"), ("builtins", "slice"), ("builtins", "complex"),
    ("builtins", "bytearray"), ("builtins", "list"), ("builtins", "dict"), ("builtins", "tuple"),
    ("builtins", "int"), ("builtins", "float"), ("builtins", "bool"), ("builtins", "str"), ("builtins", "bytes"),
} | {("torch", n) for n in _STORAGES} | {("torch.storage", "UntypedStorage"), ("torch.storage", "TypedStorage")}


def _placeholder(module, name):
    return type(name, (), {"__module__": module, "__init__": lambda self, *a, **k: None,
                           "__setstate__": lambda self, state: None, "__reduce__": lambda self: (str, ("",))})


class _RestrictedUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "easydict" and name == "EasyDict":
            return EasyDict
        if (module, name) in _SAFE:
            return super().find_class(module, name)
        if module == "torch" and name.endswith("dtype"):
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
