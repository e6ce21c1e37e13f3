"""Host-side containers of the tracking API: `PointCloud` and `Box` with the attribute names the reference's frame loop
uses (datasets/data_classes.py:11-125, :128-257), so sequences built for the reference (`{"pc": PointCloud, "3d_bbox": Box}`)
can be handed to `evaluate_one_sequence`.  The orientation is a 3x3 rotation matrix (`rotation_matrix`) instead of a
pyquaternion object; the geometry itself lives in open3dsot_b200/tracking/boxes.py as tensor math."""
import numpy as np
import torch

from ..tracking import boxes as bx


class PointCloud:
    def __init__(self, points):
        points = np.asarray(points)
        self.points = points[0:3, :] if points.shape[0] > 3 else points     # (3, N), as upstream

    def nbr_points(self):
        return self.points.shape[1]


class Box:
    def __init__(self, center, size, orientation=None, yaw_degrees=0.0):
        self.center = np.asarray(center, dtype=np.float64)
        self.wlh = np.asarray(size, dtype=np.float64)
        if orientation is None:
            a = np.deg2rad(yaw_degrees)
            orientation = np.array([[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]])
        self.rotation_matrix = np.asarray(getattr(orientation, "rotation_matrix", orientation), dtype=np.float64)

    def to_tensor(self, device=None, dtype=torch.float32):
        return bx.Box(torch.as_tensor(self.center, dtype=dtype, device=device), torch.as_tensor(self.wlh, dtype=dtype, device=device),
                      torch.as_tensor(self.rotation_matrix, dtype=dtype, device=device))

    @classmethod
    def from_tensor(cls, box: bx.Box):
        return cls(box.center.detach().cpu().double().numpy(), box.wlh.detach().cpu().double().numpy(),
                   box.rot.detach().cpu().double().numpy())

    def corners(self, wlh_factor=1.0):
        return bx.corners(self.to_tensor(dtype=torch.float64), wlh_factor).numpy().T            # (3, 8), as upstream

    def bottom_corners(self):
        return self.corners()[:, [2, 3, 7, 6]]
