"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the B=1 tracking-inference path of SURVEY.md §8(f) rank 2: the geometry the
reference's frame loop runs on the host between two model calls.

PINNED (round 2) for the geometry: tests/golden/make_golden_tracking.py runs the reference's own, unmodified
datasets/points_utils.py / data_classes.py / sampler.py (with numpy stand-ins for pyquaternion and nuscenes-devkit's
points_in_box, tests/golden/_ref_shims.py) and commits its outputs; tests/test_tracking_golden.py holds every geometry function
here and the whole `siamese_processing` to them at 1e-9 / 1e-5 on the recorded random draws.  Still unpinned: the metrics
(utils/metrics.py needs shapely / torchmetrics) and the motion-centric `motion_processing`.
This file restates the reference arithmetic with plain numpy, one function per reference function, each
citing the lines it follows; orientations are carried as 3x3 rotation matrices (pyquaternion is only used upstream to
compose and invert rotations: `Quaternion(matrix=M)` == M, `.inverse` == M.T, `q1 * q2` == M1 @ M2,
`Quaternion(axis=[0,0,1], degrees=a)` == rotz(a)), and polygon clipping replaces shapely for the two convex
quadrilaterals of a bird's-eye-view IoU.

Only tests/ may import this module."""
import copy

import numpy as np


def rotz(angle, degrees=True):
    a = np.deg2rad(angle) if degrees else angle
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


class Box:
    """center (3,), wlh (3,), rot (3,3).  datasets/data_classes.py:128-257 with the quaternion replaced by its matrix."""

    def __init__(self, center, wlh, rot):
        self.center = np.array(center, dtype=np.float64)
        self.wlh = np.array(wlh, dtype=np.float64)
        self.rot = np.array(rot, dtype=np.float64)

    def translate(self, x):                       # data_classes.py:205-211
        self.center = self.center + np.asarray(x, dtype=np.float64)

    def rotate(self, m):                          # data_classes.py:213-221  (q * orientation, centre rotated too)
        self.center = m @ self.center
        self.rot = m @ self.rot

    def corners(self, wlh_factor=1.0):            # data_classes.py:229-252: x forward (l), y left (w), z up (h)
        w, l, h = self.wlh * wlh_factor
        x = l / 2 * np.array([1, 1, 1, 1, -1, -1, -1, -1])
        y = w / 2 * np.array([1, -1, -1, 1, 1, -1, -1, 1])
        z = h / 2 * np.array([1, 1, -1, -1, 1, 1, -1, -1])
        return self.rot @ np.vstack((x, y, z)) + self.center[:, None]

    def bottom_corners(self):                     # data_classes.py:254-257
        return self.corners()[:, [2, 3, 7, 6]]


def regularize_pc(points, sample_size, seed=None):
    """points_utils.py:24-40 (numpy Generator.choice when a seed is given, as the evaluation loop does with seed=1)."""
    n = points.shape[0]
    idx = None
    rng = np.random if seed is None else np.random.default_rng(seed)
    if n > 2:
        idx = rng.choice(n, size=sample_size, replace=sample_size > n) if n != sample_size else np.arange(n)
    if idx is not None:
        return points[idx, :], idx
    return np.zeros((sample_size, 3), dtype="float32"), None


def get_offset_bb(box, offset, degrees=True, use_z=False, limit_box=True, rand=None):
    """points_utils.py:43-85.  Net effect: centre += R @ (dx, dy, dz*use_z), R <- R @ rotz(angle).
    `rand` = (r_dx, r_dy): the uniform(-1, 1) numbers limit_box substitutes for out-of-range dx / dy (the reference draws
    them from numpy's global RNG; drawn here the same way when not given)."""
    if rand is None:
        rand = (np.random.uniform(-1, 1), np.random.uniform(-1, 1))
    offset = np.array(offset, dtype=np.float64)
    new_box = copy.deepcopy(box)
    rot, trans = box.rot.copy(), box.center.copy()
    new_box.translate(-trans)
    new_box.rotate(rot.T)
    if len(offset) == 3:
        use_z = False
    ang = offset[2] if len(offset) == 3 else offset[3]
    new_box.rotate(rotz(ang, degrees))
    if limit_box:
        if offset[0] > new_box.wlh[0]:
            offset[0] = rand[0]
        if offset[1] > min(new_box.wlh[1], 2):
            offset[1] = rand[1]
        if use_z and offset[2] > new_box.wlh[2]:
            offset[2] = 0
    new_box.translate(np.array([offset[0], offset[1], offset[2] if use_z else 0.0]))
    new_box.rotate(rot)
    new_box.translate(trans)
    return new_box


def crop_pc_axis_aligned(points, box, offset=0, scale=1.0):
    """points_utils.py:147-173 on a (3, N) array; strict inequalities."""
    tmp = copy.deepcopy(box)
    tmp.wlh = tmp.wlh * scale
    c = tmp.corners()
    maxi, mini = c.max(1) + offset, c.min(1) - offset
    keep = np.ones(points.shape[1], dtype=bool)
    for a in range(3):
        keep &= (points[a] > mini[a]) & (points[a] < maxi[a])
    return points[:, keep], keep


def crop_and_center_pc(points, box, offset=0, scale=1.0):
    """points_utils.py:102-124: coarse world-frame crop (4x scale, 2x offset), move into the box frame, exact crop."""
    pts, _ = crop_pc_axis_aligned(points, box, offset=2 * offset, scale=4 * scale)
    new_box = copy.deepcopy(box)
    rot_t, trans = box.rot.T.copy(), -box.center
    pts = rot_t @ (pts + trans[:, None])
    new_box.translate(trans)
    new_box.rotate(rot_t)
    pts, _ = crop_pc_axis_aligned(pts, new_box, offset=offset, scale=scale)
    return pts, new_box


def get_model(pcs, boxes, offset=0, scale=1.0):
    """points_utils.py:88-100: merged canonical template; the returned box is the one of the LAST pair."""
    parts = [np.ones((3, 0))]
    new_box = None
    for pts, box in zip(pcs, boxes):
        cropped, new_box = crop_and_center_pc(pts, box, offset=offset, scale=scale)
        if cropped.shape[1] > 0:
            parts.append(cropped)
    return np.concatenate(parts, axis=1), new_box


def generate_subwindow(points, box, scale, offset=2):
    """points_utils.py:223-254, oriented=True: points in the frame of `box`, cropped to its scaled + padded extent."""
    tmp = copy.deepcopy(box)
    rot_t, trans = box.rot.T.copy(), -box.center
    pts = rot_t @ (points + trans[:, None])
    tmp.translate(trans)
    tmp.rotate(rot_t)
    out, _ = crop_pc_axis_aligned(pts, tmp, scale=scale, offset=offset)
    return out


def get_point_to_box_distance(points, box, wlh_factor=1.0):
    """points_utils.py:127-144: (N, 9) distances to the centre and the 8 corners (the BoxCloud)."""
    ref = np.concatenate([box.center.reshape(3, 1), box.corners(wlh_factor)], axis=1).T     # (9, 3)
    return np.sqrt(((points[:, None, :] - ref[None, :, :]) ** 2).sum(-1))


# ---- metrics (utils/metrics.py) ---------------------------------------------------------------------------------
def _clip(subject, clipper):
    """Sutherland-Hodgman: convex `subject` polygon clipped by convex `clipper`, both counter-clockwise (k, 2)."""
    out = [tuple(p) for p in subject]
    for i in range(len(clipper)):
        a, b = clipper[i], clipper[(i + 1) % len(clipper)]
        inp, out = out, []
        if not inp:
            break

        def inside(p):
            return (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0]) >= 0

        def cross(p, q):
            d1, d2 = (p[0] - q[0], p[1] - q[1]), (a[0] - b[0], a[1] - b[1])
            den = d1[0] * d2[1] - d1[1] * d2[0]
            n1, n2 = p[0] * q[1] - p[1] * q[0], a[0] * b[1] - a[1] * b[0]
            return ((n1 * d2[0] - d1[0] * n2) / den, (n1 * d2[1] - d1[1] * n2) / den)
        s = inp[-1]
        for e in inp:
            if inside(e):
                if not inside(s):
                    out.append(cross(s, e))
                out.append(e)
            elif inside(s):
                out.append(cross(s, e))
            s = e
    return np.array(out).reshape(-1, 2)


def _area(poly):
    if len(poly) < 3:
        return 0.0
    x, y = poly[:, 0], poly[:, 1]
    return 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))


def _ccw(poly):
    x, y = poly[:, 0], poly[:, 1]
    return poly if (np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1))) > 0 else poly[::-1]


def box_to_poly(box, up_axis=(0, -1, 0)):
    """utils/metrics.py:37-47."""
    if up_axis[1] != 0:
        return box.corners()[[0, 2]].T[[0, 1, 5, 4]]
    return box.bottom_corners().T[:, :2]      # shapely ignores the third coordinate for areas


def estimate_overlap(box_a, box_b, dim=2, up_axis=(0, -1, 0)):
    """utils/metrics.py:50-74."""
    pa, pb = _ccw(box_to_poly(box_a, up_axis)), _ccw(box_to_poly(box_b, up_axis))
    inter = _area(_clip(pa, pb))
    union = _area(pa) + _area(pb) - inter
    if dim == 2:
        return inter / union if union > 0 else 0.0      # degenerate footprint: shapely raises, the reference returns 0
    up = np.array(up_axis) != 0
    up_max = min(box_a.center[up], box_b.center[up])
    up_min = max(box_a.center[up] - box_a.wlh[2], box_b.center[up] - box_b.wlh[2])
    inter_vol = inter * max(0, up_max[0] - up_min[0])
    va, vb = np.prod(box_a.wlh), np.prod(box_b.wlh)
    return inter_vol / (va + vb - inter_vol)


def estimate_accuracy(box_a, box_b, dim=3, up_axis=(0, -1, 0)):
    """utils/metrics.py:27-34."""
    if dim == 3:
        return np.linalg.norm(box_a.center - box_b.center)
    keep = np.array(up_axis) != 0
    return np.linalg.norm(box_a.center[keep] - box_b.center[keep])


def success(overlaps, n=21, max_overlap=1.0):
    """utils/metrics.py:104-128: area under the success curve, in percent."""
    xs = np.linspace(0, max_overlap, n)
    o = np.asarray(overlaps, dtype=np.float64)
    if o.size == 0:
        return 0.0
    ys = np.array([(o >= t).mean() for t in xs])
    return float(np.trapezoid(ys, xs) * 100 / max_overlap)


def precision(accs, n=21, max_accuracy=2.0):
    """utils/metrics.py:77-101."""
    xs = np.linspace(0, max_accuracy, n)
    a = np.asarray(accs, dtype=np.float64)
    if a.size == 0:
        return 0.0
    ys = np.array([(a <= t).mean() for t in xs])
    return float(np.trapezoid(ys, xs) * 100 / max_accuracy)


# ---- training batch construction (datasets/sampler.py) -------------------------------------------------------------
def transform_box(box, ref):
    """points_utils.py:257-262."""
    b = copy.deepcopy(box)
    b.translate(-ref.center)
    b.rotate(ref.rot.T)
    return b


def get_in_box_mask(points, box):
    """points_utils.py:273-300 on a (3, N) array."""
    tmp = copy.deepcopy(box)
    rot_t, trans = box.rot.T.copy(), -box.center
    pts = rot_t @ (points + trans[:, None])
    tmp.translate(trans)
    tmp.rotate(rot_t)
    c = tmp.corners()
    maxi, mini = c.max(1), c.min(1)
    keep = np.ones(points.shape[1], dtype=bool)
    for a in range(3):
        keep &= (pts[a] > mini[a]) & (pts[a] < maxi[a])
    return keep


def siamese_processing(first, template, search, candidate_id, cfg, template_offset, search_offset, idx_t=None, idx_s=None,
                       limit_rand_t=None, limit_rand_s=None, aug_search=None):
    """datasets/sampler.py:16-79 with the random draws passed in: `template_offset` = the uniform(-0.3, 0.3) triple,
    `search_offset` = the KalmanFiltering sample; `idx_t` / `idx_s` = regularize_pc's index draws.
    Frames are (points (3, N), Box).  Returns the data dict plus the survivor counts."""
    (first_pc, first_box), (t_pc, t_box), (s_pc, s_box) = first, template, search
    if aug_search is not None:                                     # search_transform (sampler.py:34-35)
        s_pc, s_box = apply_augmentation(s_pc, s_box, *aug_search)
    deg = 5 if cfg["degrees"] else np.deg2rad(5)
    if candidate_id == 0:
        off_t = np.zeros(3)
    else:
        off_t = np.array(template_offset, dtype=np.float64)
        off_t[2] = off_t[2] * deg
    t_box = get_offset_bb(t_box, off_t, limit_box=cfg["data_limit_box"], degrees=cfg["degrees"], rand=limit_rand_t)
    model_pc, model_box = get_model([first_pc, t_pc], [first_box, t_box], scale=cfg["model_bb_scale"], offset=cfg["model_bb_offset"])
    if candidate_id == 0 and cfg.get("num_candidates", 1) > 1:
        off_s = np.zeros(3)
    else:
        off_s = np.array(search_offset, dtype=np.float64)
    sample_bb = get_offset_bb(s_box, off_s, limit_box=cfg["data_limit_box"], degrees=cfg["degrees"], rand=limit_rand_s)
    crop = generate_subwindow(s_pc, sample_bb, scale=cfg["search_bb_scale"], offset=cfg["search_bb_offset"])
    s_box = transform_box(s_box, sample_bb)
    seg = get_in_box_mask(crop, s_box).astype(int)
    reg = [s_box.center[0], s_box.center[1], s_box.center[2], -off_s[2]]
    tp = model_pc.T[idx_t] if idx_t is not None else regularize_pc(model_pc.T, cfg["template_size"])[0]
    if idx_s is None:
        sp, idx_s = regularize_pc(crop.T, cfg["search_size"])
    else:
        sp = crop.T[idx_s]
    out = {"template_points": tp.astype("float32"), "search_points": sp.astype("float32"), "box_label": np.array(reg, dtype="float32"),
           "bbox_size": s_box.wlh, "seg_label": seg[idx_s].astype("float32"), "n_template": model_pc.shape[1], "n_search": crop.shape[1],
           "_model_pc": model_pc, "_model_box": model_box, "_crop": crop, "_seg": seg, "_search_box": s_box}
    if cfg.get("box_aware", False):
        out["points2cc_dist_t"] = get_point_to_box_distance(tp, model_box).astype("float32")
        out["points2cc_dist_s"] = get_point_to_box_distance(sp, s_box).astype("float32")
    return out


def motion_build_input(prev_pts, this_pts, ref_box, cfg, frame_id):
    """models/base_model.py:255-303 (MotionBaseModel.build_input_dict) on (3, N) scans; regularize_pc with seed=1.
    nuscenes' points_in_box(box, pts, 1.25) is restated as the inclusive test in the box frame (its three projections
    0 <= v.e <= e.e are exactly -half <= local <= half)."""
    n = cfg["point_sample_size"]
    prev_crop = generate_subwindow(prev_pts, ref_box, cfg["bb_scale"], cfg["bb_offset"])
    this_crop = generate_subwindow(this_pts, ref_box, cfg["bb_scale"], cfg["bb_offset"])
    canon = transform_box(ref_box, ref_box)
    pp, _ = regularize_pc(prev_crop.T, n, seed=1)
    tp, _ = regularize_pc(this_crop.T, n, seed=1)
    half = np.array([canon.wlh[1], canon.wlh[0], canon.wlh[2]]) * 1.25 / 2
    mask_prev = (np.abs(pp) <= half).all(1).astype(float)
    if frame_id != 1:
        mask_prev = np.where(mask_prev == 0, 0.2, 0.8)
    mask_this = np.full(mask_prev.shape, 0.5)
    prev = np.concatenate([pp, np.full((n, 1), 0.0), mask_prev[:, None]], -1)
    this = np.concatenate([tp, np.full((n, 1), 0.1), mask_this[:, None]], -1)
    stack = np.concatenate([prev, this], 0)
    out = {"points": stack[None].astype("float32")}
    if cfg.get("box_aware", False):
        bc = get_point_to_box_distance(stack[:n, :3], canon)
        out["candidate_bc"] = np.concatenate([bc, np.zeros_like(bc)], 0)[None].astype("float32")
    return out


def _yaw(rot, degrees):
    a = np.arctan2(rot[1, 0], rot[0, 0])
    return np.rad2deg(a) if degrees else a


def _in_box_inclusive(points, box, factor):
    """nuscenes geometry_utils.points_in_box on (N, 3) points."""
    local = (points - box.center) @ box.rot
    half = np.array([box.wlh[1], box.wlh[0], box.wlh[2]]) * factor / 2
    return (np.abs(local) <= half).all(1)


def motion_processing(prev, this, candidate_id, cfg, offset, idx_prev=None, idx_this=None, limit_rand=None, aug_prev=None,
                      aug_this=None):
    """datasets/sampler.py:82-181 with the random draws passed in.  Frames are (points (3, N), Box); `aug_prev` / `aug_this`
    = (trans, rot_deg, flip_x, flip_y) of the optional augmentation transform (:102-105)."""
    (prev_pc, prev_box), (this_pc, this_box) = prev, this
    n = cfg["point_sample_size"]
    deg = 5 if cfg["degrees"] else np.deg2rad(5)
    n_target = int(_in_box_inclusive(prev_pc.T, prev_box, 1.0).sum())
    if aug_prev is not None:
        prev_pc, prev_box = apply_augmentation(prev_pc, prev_box, *aug_prev)
    if aug_this is not None:
        this_pc, this_box = apply_augmentation(this_pc, this_box, *aug_this)
    if candidate_id == 0:
        off = np.zeros(3)
    else:
        off = np.array(offset, dtype=np.float64)
        off[2] = off[2] * deg
    ref = get_offset_bb(prev_box, off, limit_box=cfg["data_limit_box"], degrees=cfg["degrees"], rand=limit_rand)
    prev_crop = generate_subwindow(prev_pc, ref, cfg["bb_scale"], cfg["bb_offset"])
    this_crop = generate_subwindow(this_pc, ref, cfg["bb_scale"], cfg["bb_offset"])
    this_b, prev_b, ref_b = transform_box(this_box, ref), transform_box(prev_box, ref), transform_box(ref, ref)
    motion_b = transform_box(this_b, prev_b)
    pp = prev_crop.T[idx_prev] if idx_prev is not None else regularize_pc(prev_crop.T, n)[0]
    tp = this_crop.T[idx_this] if idx_this is not None else regularize_pc(this_crop.T, n)[0]
    seg_this = _in_box_inclusive(tp, this_b, 1.25).astype(int)
    seg_prev = _in_box_inclusive(pp, prev_b, 1.25).astype(int)
    mask_prev = _in_box_inclusive(pp, ref_b, 1.25).astype(float)
    if candidate_id != 0:
        mask_prev = np.where(mask_prev == 0, 0.2, 0.8)
    stack = np.concatenate([np.concatenate([pp, np.full((n, 1), 0.0), mask_prev[:, None]], -1),
                            np.concatenate([tp, np.full((n, 1), 0.1), np.full((n, 1), 0.5)], -1)], 0)
    lab = lambda b: np.append(b.center, _yaw(b.rot, cfg["degrees"])).astype("float32")
    out = {"points": stack.astype("float32"), "box_label": lab(this_b), "box_label_prev": lab(prev_b), "motion_label": lab(motion_b),
           "motion_state_label": int(np.sqrt(((this_b.center - prev_b.center) ** 2).sum()) > cfg["motion_threshold"]),
           "bbox_size": this_b.wlh, "seg_label": np.hstack([seg_prev, seg_this]).astype(int),
           "n_prev": prev_crop.shape[1], "n_this": this_crop.shape[1], "n_target": n_target,
           "_prev_crop": prev_crop, "_this_crop": this_crop, "_boxes": (this_b, prev_b, ref_b)}
    if cfg.get("box_aware", False):
        c = get_point_to_box_distance(pp, ref_b)
        out.update({"prev_bc": get_point_to_box_distance(pp, prev_b).astype("float32"),
                    "this_bc": get_point_to_box_distance(tp, this_b).astype("float32"),
                    "candidate_bc": np.concatenate([c, np.zeros_like(c)], 0).astype("float32")})
    return out


def apply_augmentation(points, box, trans, rot_deg, flip_x, flip_y, wlh_factor=1.25):
    """points_utils.py:303-362 (apply_transform + apply_augmentation) with the random draws passed in; points (3, N)."""
    inside = _in_box_inclusive(points.T, box, wlh_factor)
    rot, c = box.rot.copy(), box.center.copy()
    new_box = copy.deepcopy(box)
    pts = points[:, inside].copy()
    pts = rot.T @ (pts - c[:, None])
    new_box.translate(-c)
    new_box.rotate(rot.T)
    if flip_x:
        pts[0] = -pts[0]
        new_box.rotate(rotz(180.0))
    if flip_y:
        pts[1] = -pts[1]
    q = rotz(rot_deg)
    new_box.rotate(q)
    pts = q @ pts
    new_box.translate(np.asarray(trans, dtype=np.float64))
    pts = pts + np.asarray(trans, dtype=np.float64)[:, None]
    new_box.rotate(rot)
    pts = rot @ pts
    new_box.translate(c)
    pts = pts + c[:, None]
    out = points.copy()
    out[:, inside] = pts
    return out, new_box
