"""Drop-in for `pointnet2_ops._ext` (the module the reference imports at
pointnet2/utils/pointnet2_utils.py:17): the same nine functions, same argument order, same results,
backed by libo3d_b200.so.  `import open3dsot_b200._ext as _ext` is the one-line change a maintainer
makes to run the reference's unmodified Python on these kernels."""
from .ops import (ball_query, furthest_point_sampling, gather_points, gather_points_grad, group_points,  # noqa: F401
                  group_points_grad, three_interpolate, three_interpolate_grad, three_nn)

__all__ = ["furthest_point_sampling", "gather_points", "gather_points_grad", "ball_query", "group_points",
           "group_points_grad", "three_nn", "three_interpolate", "three_interpolate_grad"]
