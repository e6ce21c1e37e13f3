"""ORACLE (test infrastructure): torch-CPU wrappers over the C restatement of `pointnet2_ops._ext`.

Signatures follow the upstream pybind entry points as the reference calls them
(pointnet2/utils/pointnet2_utils.py:56,92,98,125,162,184,217,237,268): note `ball_query`
takes (new_xyz, xyz, radius, nsample) — the reverse of the Python-level wrapper.
"""
import ctypes
import os

import torch

from . import build as _build

_lib = None
FMA_MODE = 1  # upstream is compiled with nvcc --fmad=true (see pointnet2_ops_ref.c header)


def lib():
    global _lib
    if _lib is None:
        path = _build.OUT
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(_build.SRC):
            path = _build.build()
        _lib = ctypes.CDLL(path)
    return _lib


def _f(t):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu", "oracle: CPU contiguous f32"
    return ctypes.c_void_p(t.data_ptr())


def _i(t):
    assert t.dtype == torch.int32 and t.is_contiguous() and t.device.type == "cpu", "oracle: CPU contiguous i32"
    return ctypes.c_void_p(t.data_ptr())


def opt_n_threads(n: int) -> int:
    return lib().o3d_ref_opt_n_threads(int(n))


def furthest_point_sampling(xyz, npoint):
    B, N, _ = xyz.shape
    out = torch.zeros(B, npoint, dtype=torch.int32)
    lib().o3d_ref_fps(_f(xyz), B, N, int(npoint), _i(out), FMA_MODE)
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    out = torch.zeros(B, M, nsample, dtype=torch.int32)
    lib().o3d_ref_ball_query(_f(new_xyz), _f(xyz), B, N, M, ctypes.c_float(radius), int(nsample), _i(out), FMA_MODE)
    return out


def gather_points(features, idx):
    B, C, N = features.shape
    M = idx.shape[1]
    out = torch.zeros(B, C, M)
    lib().o3d_ref_gather(_f(features), _i(idx), B, C, N, M, _f(out))
    return out


def gather_points_grad(grad_out, idx, N):
    B, C, M = grad_out.shape
    out = torch.zeros(B, C, N)
    lib().o3d_ref_gather_grad(_f(grad_out), _i(idx), B, C, N, M, _f(out))
    return out


def group_points(features, idx):
    B, C, N = features.shape
    _, M, S = idx.shape
    out = torch.zeros(B, C, M, S)
    lib().o3d_ref_group(_f(features), _i(idx), B, C, N, M, S, _f(out))
    return out


def group_points_grad(grad_out, idx, N):
    B, C, M, S = grad_out.shape
    out = torch.zeros(B, C, N)
    lib().o3d_ref_group_grad(_f(grad_out), _i(idx), B, C, N, M, S, _f(out))
    return out


def three_nn(unknown, known):
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = torch.zeros(B, n, 3)
    idx = torch.zeros(B, n, 3, dtype=torch.int32)
    lib().o3d_ref_three_nn(_f(unknown), _f(known), B, n, m, _f(dist2), _i(idx), FMA_MODE)
    return dist2, idx


def three_interpolate(features, idx, weight):
    B, c, m = features.shape
    n = idx.shape[1]
    out = torch.zeros(B, c, n)
    lib().o3d_ref_three_interpolate(_f(features), _i(idx), _f(weight), B, c, m, n, _f(out), FMA_MODE)
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    B, c, n = grad_out.shape
    out = torch.zeros(B, c, m)
    lib().o3d_ref_three_interpolate_grad(_f(grad_out), _i(idx), _f(weight), B, c, n, m, _f(out))
    return out
