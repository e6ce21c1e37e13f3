"""Batched tensor box transforms used inside M2-Track's forward pass.

Mirror of the torch part of datasets/points_utils.py (:364-451): rotz_batch_tensor, get_offset_points_tensor,
get_offset_box_tensor, remove_transform_points_tensor.  (The numpy / pyquaternion crop-and-label code of that file is
CPU data preparation and out of scope — SURVEY.md §2.1 row 12.)  Functional (no in-place edits of the inputs)."""
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
