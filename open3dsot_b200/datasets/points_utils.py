"""Batched tensor box transforms used inside M2-Track's forward pass.

Mirror of the torch part of datasets/points_utils.py (:364-451): rotz_batch_tensor, get_offset_points_tensor,
get_offset_box_tensor, remove_transform_points_tensor, functional (no in-place edits of the inputs).

The second half of this module gives the file's host-level geometry (:24-300: regularize_pc, getOffsetBB, getModel,
cropAndCenterPC, crop_pc_axis_aligned, generate_subwindow, get_point_to_box_distance, transform_box, transform_pc,
get_in_box_mask) under the reference's names and call signatures, for code that was written against them: arguments and
results are the host containers of datasets/data_classes.py (`PointCloud` (3, N), `Box` with a rotation matrix), the
arithmetic is the tensor geometry of open3dsot_b200/tracking/boxes.py (checked against the numpy restatement in
tests/test_tracking_host.py)."""
import numpy as np
import torch


def rotz_batch_tensor(t):
    """(...,) angles -> (...,3,3) rotation matrices about z."""
    c, s = torch.cos(t), torch.sin(t)
    z, o = torch.zeros_like(t), torch.ones_like(t)
    return torch.stack([torch.stack([c, -s, z], -1), torch.stack([s, c, z], -1), torch.stack([z, z, o], -1)], -2)


def roty_batch_tensor(t):
    c, s = torch.cos(t), torch.sin(t)
    z, o = torch.zeros_like(t), torch.ones_like(t)
    return torch.stack([torch.stack([c, z, s], -1), torch.stack([z, o, z], -1), torch.stack([-s, z, c], -1)], -2)


def get_offset_points_tensor(points, ref_box_params, offset_box_params):
    """Move points (B,N,3) by an offset (B,4) expressed in the frame of ref box (B,4): world -> box frame -> offset -> world."""
    ref_center, ref_rot = ref_box_params[:, :3], ref_box_params[:, -1]
    off_center, off_rot = offset_box_params[:, :3], offset_box_params[:, -1]
    rot_mat = rotz_batch_tensor(-ref_rot)
    pts = torch.matmul(points - ref_center[:, None, :], rot_mat.transpose(1, 2))
    pts = torch.matmul(pts, rotz_batch_tensor(off_rot).transpose(1, 2)) + off_center[:, None, :]
    return torch.matmul(pts, rot_mat) + ref_center[:, None, :]


def get_offset_box_tensor(ref_box_params, offset_box_params):
    """Compose a box (B,4) = [x,y,z,theta] with an offset (B,4) given in the box frame."""
    ref_center, ref_rot = ref_box_params[:, :3], ref_box_params[:, -1]
    off_center, off_rot = offset_box_params[:, :3], offset_box_params[:, -1]
    new_center = torch.matmul(rotz_batch_tensor(ref_rot), off_center[..., None]).squeeze(-1) + ref_center
    return torch.cat([new_center, (ref_rot + off_rot)[:, None]], dim=-1)


def remove_transform_points_tensor(points, ref_box_params):
    """Express points (B,N,3) in the frame of the box (B,4)."""
    ref_center, ref_rot = ref_box_params[:, :3], ref_box_params[:, -1]
    return torch.matmul(points - ref_center[:, None, :], rotz_batch_tensor(-ref_rot).transpose(1, 2))


# ---- the reference's host-level names (datasets/points_utils.py:24-300) ------------------------------------------
def _bx():
    from ..tracking import boxes
    return boxes


def _pts(pc):
    from .data_classes import PointCloud
    return torch.as_tensor(np.ascontiguousarray((pc.points if isinstance(pc, PointCloud) else np.asarray(pc).T).T), dtype=torch.float64)


def _pc(t):
    from .data_classes import PointCloud
    return PointCloud(t.numpy().T.copy())


def _tb(box):
    return box.to_tensor(dtype=torch.float64)


def regularize_pc(points, sample_size, seed=None):
    """:24-40.  points (N, 3) numpy -> (sample_size, 3), index array (or None when fewer than 3 points)."""
    n = points.shape[0]
    if n <= 2:
        return np.zeros((sample_size, 3), dtype='float32'), None
    rng = np.random if seed is None else np.random.default_rng(seed)
    idx = rng.choice(n, size=sample_size, replace=sample_size > n) if n != sample_size else np.arange(n)
    return points[idx, :], idx


def getOffsetBB(box, offset, degrees=True, use_z=False, limit_box=True, inplace=False):
    """:43-85 (the out-of-range replacements of limit_box come from numpy's global RNG, as upstream)."""
    from .data_classes import Box
    off = torch.as_tensor(np.asarray(offset, dtype=np.float64))
    rand = torch.as_tensor(np.random.uniform(-1, 1, 2)) if limit_box else None
    new = Box.from_tensor(_bx().offset_box(_tb(box), off, degrees=degrees, use_z=use_z, limit_box=limit_box, rand=rand))
    if inplace:
        box.center, box.wlh, box.rotation_matrix = new.center, new.wlh, new.rotation_matrix
        return box
    return new


def crop_pc_axis_aligned(PC, box, offset=0, scale=1.0, return_mask=False):
    """:147-173."""
    pts = _pts(PC)
    keep = _bx().axis_aligned_mask(pts, _tb(box), offset=offset, scale=scale)
    out = _pc(pts[keep])
    return (out, keep.numpy()) if return_mask else out


def cropAndCenterPC(PC, box, offset=0, scale=1.0, normalize=False):
    """:102-124.  Returns (points in the box frame inside the scaled box, canonical box)."""
    from .data_classes import Box
    local, keep, canon = _bx().crop_and_center(_pts(PC), _tb(box), offset=offset, scale=scale)
    pts = local[keep]
    if normalize:
        pts = pts / torch.as_tensor([box.wlh[1], box.wlh[0], box.wlh[2]], dtype=pts.dtype)
    return _pc(pts), Box.from_tensor(canon)


def getModel(PCs, boxes, offset=0, scale=1.0, normalize=False):
    """:88-100: merged canonical crops; the box is that of the last pair."""
    from .data_classes import PointCloud
    if len(PCs) == 0:
        return PointCloud(np.ones((3, 0)))
    parts, new_box = [np.ones((3, 0))], None
    for PC, box in zip(PCs, boxes):
        cropped, new_box = cropAndCenterPC(PC, box, offset=offset, scale=scale, normalize=normalize)
        if cropped.nbr_points() > 0:
            parts.append(cropped.points)
    return PointCloud(np.concatenate(parts, axis=1)), new_box


def generate_subwindow(pc, sample_bb, scale, offset=2, oriented=True):
    """:223-254 (both variants return the crop in the frame of `sample_bb`)."""
    b = _bx()
    pts, box = _pts(pc), _tb(sample_bb)
    if oriented:
        local, keep = b.subwindow(pts, box, scale=scale, offset=offset)
    else:
        keep = b.axis_aligned_mask(pts, box, offset=offset, scale=scale)
        local = b.to_box_frame(pts, box)
    return _pc(local[keep])


def get_point_to_box_distance(pc, box, wlh_factor=1.0):
    """:127-144: (N, 9) distances to the centre and the eight corners."""
    return _bx().point_to_box_distance(_pts(pc), _tb(box), wlh_factor).numpy()


def transform_box(box, ref_box, inplace=False):
    """:257-262."""
    from .data_classes import Box
    b, r = _tb(box), _tb(ref_box)
    new = Box(((b.center - r.center) @ r.rot).numpy(), box.wlh, (r.rot.t() @ b.rot).numpy())
    if inplace:
        box.center, box.rotation_matrix = new.center, new.rotation_matrix
        return box
    return new


def transform_pc(pc, ref_box, inplace=False):
    """:265-270."""
    out = _pc(_bx().to_box_frame(_pts(pc), _tb(ref_box)))
    if inplace:
        pc.points = out.points
        return pc
    return out


def get_in_box_mask(PC, box):
    """:273-300: strictly inside the oriented box."""
    b = _tb(box)
    local = _bx().to_box_frame(_pts(PC), b)
    half = torch.stack([b.wlh[1], b.wlh[0], b.wlh[2]]) / 2
    return (local.abs() < half).all(-1).numpy()
