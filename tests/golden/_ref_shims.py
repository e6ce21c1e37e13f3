"""Stand-ins for the reference's absent host-side dependencies, used ONLY by the golden-vector generators (which import the
reference's own, unmodified Python from /root/reference).

* `Quaternion` — numpy restatement of the part of `pyquaternion` (un-pinned in requirement.txt; published algorithm) that
  datasets/data_classes.py, datasets/points_utils.py and datasets/kitti.py touch: construction from a rotation matrix, from
  axis + angle, from four elements; `rotation_matrix`, `inverse`, `*`, `elements`, `axis`, `radians`, `degrees`.
  A quaternion and its negative are the same rotation; nothing the reference computes on this path depends on the sign.
* `points_in_box` — nuscenes-devkit `nuscenes.utils.geometry_utils.points_in_box` (published algorithm: the point is projected
  on the three box edges that meet at corner 0 and must fall inside each edge's extent, inclusive)."""
import numpy as np


class Quaternion:
    def __init__(self, *args, **kw):
        if "matrix" in kw:
            self.q = self._from_matrix(np.asarray(kw["matrix"], dtype=np.float64))
        elif "axis" in kw:
            ang = kw["radians"] if "radians" in kw else (np.deg2rad(kw["degrees"]) if "degrees" in kw else kw.get("angle", 0.0))
            ax = np.asarray(kw["axis"], dtype=np.float64)
            ax = ax / np.linalg.norm(ax)
            self.q = np.concatenate([[np.cos(ang / 2.0)], ax * np.sin(ang / 2.0)])
        elif len(args) == 0:
            self.q = np.array([1.0, 0.0, 0.0, 0.0])
        elif len(args) == 1 and isinstance(args[0], Quaternion):
            self.q = args[0].q.copy()
        elif len(args) == 1:
            self.q = np.asarray(args[0], dtype=np.float64).reshape(4).copy()
        else:
            self.q = np.asarray(args, dtype=np.float64).reshape(4).copy()

    @staticmethod
    def _from_matrix(m):
        m = m[:3, :3]
        assert np.allclose(m @ m.T, np.eye(3), atol=1e-6) and np.isclose(np.linalg.det(m), 1.0, atol=1e-6), "not a rotation"
        t = np.trace(m)
        if t > 0:
            s = np.sqrt(t + 1.0) * 2
            q = [0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s]
        elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
            s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
            q = [(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s]
        elif m[1, 1] > m[2, 2]:
            s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
            q = [(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s]
        else:
            s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
            q = [(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s]
        return np.asarray(q, dtype=np.float64)

    @property
    def elements(self):
        return self.q

    @property
    def rotation_matrix(self):
        w, x, y, z = self.q / np.linalg.norm(self.q)
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])

    @property
    def inverse(self):
        n = float(self.q @ self.q)
        return Quaternion(np.array([self.q[0], -self.q[1], -self.q[2], -self.q[3]]) / n)

    def __mul__(self, o):
        a, b = self.q, o.q
        return Quaternion(np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                                    a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                                    a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                                    a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]]))

    @property
    def radians(self):
        q = self.q / np.linalg.norm(self.q)
        ang = 2.0 * np.arctan2(np.linalg.norm(q[1:]), q[0])
        return float(((ang + np.pi) % (2 * np.pi)) - np.pi)

    @property
    def degrees(self):
        return float(np.rad2deg(self.radians))

    @property
    def axis(self):
        n = np.linalg.norm(self.q[1:])
        return self.q[1:] / n if n > 1e-14 else np.array([0.0, 0.0, 0.0])


def points_in_box(box, points, wlh_factor=1.0):
    corners = box.corners(wlh_factor=wlh_factor)
    p1, p_x, p_y, p_z = corners[:, 0], corners[:, 4], corners[:, 1], corners[:, 3]
    i, j, k = p_x - p1, p_y - p1, p_z - p1
    v = points - p1.reshape((-1, 1))
    iv, jv, kv = np.dot(i, v), np.dot(j, v), np.dot(k, v)
    mask_x = np.logical_and(0 <= iv, iv <= np.dot(i, i))
    mask_y = np.logical_and(0 <= jv, jv <= np.dot(j, j))
    mask_z = np.logical_and(0 <= kv, kv <= np.dot(k, k))
    return np.logical_and(np.logical_and(mask_x, mask_y), mask_z)


def install(ref_root="/root/reference"):
    """Register the shims + inert stubs so that `datasets.data_classes`, `datasets.points_utils`, `datasets.sampler` of the
    reference import unmodified (their package `__init__` — which drags in the dataset readers — is bypassed)."""
    import os
    import sys
    import types

    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m
    from open3dsot_b200.compat.easydict import EasyDict
    mod("easydict", EasyDict=EasyDict)
    mod("pyquaternion", Quaternion=Quaternion)
    nus = mod("nuscenes")
    nus.utils = mod("nuscenes.utils")
    nus.utils.geometry_utils = mod("nuscenes.utils.geometry_utils", points_in_box=points_in_box)
    mod("pomegranate", MultivariateGaussianDistribution=object, GeneralMixtureModel=object)   # searchspace.py:2 (GMM search: unused)
    ds = mod("datasets")
    ds.__path__ = [os.path.join(ref_root, "datasets")]
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
