"""ORACLE (test infrastructure): a stand-in for `pointnet2_ops._ext` on CPU tensors.

`install()` registers it in sys.modules so the reference's unmodified
pointnet2/utils/pointnet2_utils.py:17 (`import pointnet2_ops._ext as _ext`) resolves to the
C oracle — used only by tests/golden/make_golden.py inside the authoring container.
"""
import sys
import types

from . import ops

NAMES = [
    "furthest_point_sampling", "gather_points", "gather_points_grad", "ball_query",
    "group_points", "group_points_grad", "three_nn", "three_interpolate", "three_interpolate_grad",
]


def make_module():
    ext = types.ModuleType("pointnet2_ops._ext")
    for n in NAMES:
        setattr(ext, n, getattr(ops, n))
    return ext


def install():
    pkg = types.ModuleType("pointnet2_ops")
    ext = make_module()
    pkg._ext = ext
    sys.modules["pointnet2_ops"] = pkg
    sys.modules["pointnet2_ops._ext"] = ext
    return ext
