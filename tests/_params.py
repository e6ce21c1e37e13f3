"""Deterministic, name-keyed parameter fill shared by the golden generator and the tests, so that no
weights need to be stored: the same state-dict keys always receive the same values."""
import zlib

import torch


def det_tensor(key, shape, kind, seed=0):
    g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * seed) & 0x7FFFFFFF)
    if kind == "conv_w":
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return torch.randn(shape, generator=g) * (2.0 / max(fan_in, 1)) ** 0.5
    if kind == "bn_w":  # mostly positive, a few negative gammas (exercises the max/min pooling logic)
        w = 0.5 + torch.rand(shape, generator=g)
        sign = torch.where(torch.rand(shape, generator=g) < 0.1, -1.0, 1.0)
        return w * sign
    if kind == "var":
        return 0.5 + torch.rand(shape, generator=g)
    if kind == "small":
        return 0.1 * torch.randn(shape, generator=g)
    raise ValueError(kind)


def det_state_dict(reference_sd, seed=0):
    """Return a new state dict with the keys/shapes of `reference_sd` and deterministic values."""
    out = {}
    for k, v in reference_sd.items():
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros_like(v)
        elif k.endswith("running_var"):
            out[k] = det_tensor(k, v.shape, "var", seed)
        elif k.endswith("running_mean"):
            out[k] = det_tensor(k, v.shape, "small", seed)
        elif ".bn." in k and k.endswith("weight") and v.dim() == 1:
            out[k] = det_tensor(k, v.shape, "bn_w", seed)
        elif k.endswith("bias"):
            out[k] = det_tensor(k, v.shape, "small", seed)
        elif v.dim() >= 2:
            out[k] = det_tensor(k, v.shape, "conv_w", seed)
        else:  # 1-D weights of plain nn.BatchNorm1d (M2-Track nets)
            out[k] = det_tensor(k, v.shape, "bn_w", seed)
    return out
