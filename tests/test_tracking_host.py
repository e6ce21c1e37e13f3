"""B=1 tracking inference path, CPU part: the tensor geometry against the numpy restatement of
the reference (oracle/tracking_ref.py — parity unpinned, see its header), the metrics, the fixed-shape resampling, and
the frame loop's plumbing with a stand-in network."""
import numpy as np
import pytest
import torch

from oracle import tracking_ref as R
from open3dsot_b200.compat import EasyDict
from open3dsot_b200.datasets import data_classes as dc
from open3dsot_b200.datasets.synthetic import synthetic_sequence
from open3dsot_b200.models.base_model import MatchingBaseModel, regularize
from open3dsot_b200.tracking import boxes as bx
from open3dsot_b200.tracking.device_tracker import DeviceTracker
from open3dsot_b200.tracking.sampling import resample
from open3dsot_b200.utils import metrics as M


def _pair(seed):
    rng = np.random.default_rng(seed)
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    a = rng.uniform(-np.pi, np.pi)
    k = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    rot = np.eye(3) + np.sin(a) * k + (1 - np.cos(a)) * k @ k            # a general rotation, not only yaw
    if seed % 2 == 0:
        rot = R.rotz(rng.uniform(-180, 180))
    ob = R.Box(rng.normal(size=3) * 2, rng.uniform(0.5, 4.5, 3), rot)
    tb = bx.Box(torch.tensor(ob.center), torch.tensor(ob.wlh), torch.tensor(ob.rot))
    pts = rng.normal(size=(3, 4000)) * 3 + ob.center[:, None]
    return rng, ob, tb, pts


@pytest.mark.parametrize("seed", range(6))
def test_box_geometry_matches_reference_restatement(seed):
    rng, ob, tb, pts = _pair(seed)
    P = torch.tensor(pts.T)
    assert np.abs(bx.corners(tb, 1.25).numpy().T - ob.corners(1.25)).max() < 1e-12
    want = R.generate_subwindow(pts, ob, 1.25, 2)
    local, keep = bx.subwindow(P, tb, 1.25, 2.0)
    assert int(keep.sum()) == want.shape[1] and np.abs(local[keep].numpy().T - want).max() < 1e-12
    want, canon = R.crop_and_center_pc(pts, ob, offset=0.3, scale=1.25)
    local, keep, tcanon = bx.crop_and_center(P, tb, 0.3, 1.25)
    assert int(keep.sum()) == want.shape[1] and np.abs(local[keep].numpy().T - want).max() < 1e-12
    assert np.abs(canon.center).max() < 1e-12 and np.abs(canon.rot - np.eye(3)).max() < 1e-12
    assert np.abs(tcanon.wlh.numpy() - canon.wlh).max() == 0
    d = R.get_point_to_box_distance(pts.T[:200], ob)
    assert np.abs(bx.point_to_box_distance(P[:200], tb).numpy() - d).max() < 1e-12
    for off, use_z, deg in (([0.3, -0.2, 0.1, 12.0], True, True), ([0.3, -0.2, 0.1, 0.2], False, False), ([0.1, 0.2, 5.0], True, True)):
        nb = R.get_offset_bb(ob, off, degrees=deg, use_z=use_z, limit_box=False)
        tn = bx.offset_box(tb, torch.tensor(off, dtype=torch.float64), deg, use_z, False)
        assert np.abs(tn.center.numpy() - nb.center).max() < 1e-12 and np.abs(tn.rot.numpy() - nb.rot).max() < 1e-12
        assert np.abs(tn.rot.numpy() @ tn.rot.numpy().T - np.eye(3)).max() < 1e-12


def test_limit_box_replaces_out_of_range_offsets():
    _, ob, tb, _ = _pair(2)
    big = torch.tensor([ob.wlh[0] + 1.0, 0.1, ob.wlh[2] + 1.0, 3.0], dtype=torch.float64)
    rand = torch.tensor([0.25, -0.5], dtype=torch.float64)
    got = bx.offset_box(tb, big, True, True, True, rand=rand)

    want = R.get_offset_bb(ob, big.tolist(), degrees=True, use_z=True, limit_box=True, rand=(0.25, -0.5))
    assert np.abs(got.center.numpy() - want.center).max() < 1e-12


def test_get_model_merges_first_and_previous():
    rng, ob, tb, pts = _pair(4)
    ob2 = R.get_offset_bb(ob, [0.4, 0.1, 0.0, 5.0], use_z=True, limit_box=False)
    pts2 = pts + 0.05
    want, canon = R.get_model([pts, pts2], [ob, ob2], offset=0, scale=1.25)
    cfg = EasyDict(model_bb_offset=0, model_bb_scale=1.25, shape_aggregation="firstandprevious")
    m = MatchingBaseModel(cfg)
    seq = [{"pc": dc.PointCloud(pts), "3d_bbox": None}, {"pc": dc.PointCloud(pts2), "3d_bbox": None}]
    res = [dc.Box(ob.center, ob.wlh, ob.rot), dc.Box(ob2.center, ob2.wlh, ob2.rot)]
    got, tcanon = m.generate_template(seq, 2, res)
    assert got.shape[0] == want.shape[1] and np.abs(got.numpy().T - want).max() < 1e-5      # float32 points
    assert np.abs(tcanon.wlh.numpy() - canon.wlh).max() < 1e-6


def test_regularize_uses_the_reference_draw():
    pts = torch.arange(300, dtype=torch.float32).reshape(100, 3)
    for size in (64, 100, 256):
        got, idx = regularize(pts, size, seed=1)
        want, widx = R.regularize_pc(pts.numpy(), size, seed=1)
        assert np.array_equal(idx, widx) and np.array_equal(got.numpy(), want)
    got, idx = regularize(pts[:2], 64, seed=1)
    assert idx is None and float(got.abs().sum()) == 0 and got.shape == (64, 3)


def test_fixed_shape_resample():
    g = torch.Generator().manual_seed(0)
    pts = torch.randn(500, 3)
    keep = torch.zeros(500, dtype=torch.bool)
    keep[torch.randperm(500, generator=g)[:120]] = True
    out, src = resample(pts, keep, 64, g)                       # enough survivors: distinct, all kept
    assert out.shape == (64, 3) and bool(keep[src].all()) and len(set(src.tolist())) == 64
    out, src = resample(pts, keep, 256, g)                      # too few: with replacement, still only survivors
    assert out.shape == (256, 3) and bool(keep[src].all()) and len(set(src.tolist())) <= 120
    keep2 = torch.zeros(500, dtype=torch.bool)
    keep2[:2] = True
    out, _ = resample(pts, keep2, 64, g)                        # <= 2 survivors: the all-zero placeholder
    assert float(out.abs().sum()) == 0
    out, src = resample(pts[:40], keep[:40] | True, 64, g)      # fewer candidates than outputs
    assert out.shape == (64, 3) and int(src.max()) < 40
    idx = torch.tensor([3, 1, 1, 0])
    out, src = resample(pts, keep, 4, indices=idx)              # explicit draw (oracle parity)
    surv = torch.nonzero(keep)[:, 0]
    assert torch.equal(src, surv[idx]) and torch.equal(out, pts[surv[idx]])


@pytest.mark.parametrize("up_axis", [(0, 0, 1)])
def test_metrics_match_restatement_and_analytic_cases(up_axis):
    rng = np.random.default_rng(3)
    for _ in range(20):
        a = R.Box(rng.normal(size=3), rng.uniform(1, 4, 3), R.rotz(rng.uniform(-180, 180)))
        b = R.get_offset_bb(a, [rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-0.3, 0.3), rng.uniform(-40, 40)],
                            use_z=True, limit_box=False)
        ha, hb = dc.Box(a.center, a.wlh, a.rot), dc.Box(b.center, b.wlh, b.rot)
        for dim in (2, 3):
            assert abs(M.estimateOverlap(ha, hb, dim, up_axis) - R.estimate_overlap(a, b, dim, up_axis)) < 1e-9
            assert abs(M.estimateAccuracy(ha, hb, dim, up_axis) - R.estimate_accuracy(a, b, dim, up_axis)) < 1e-12
    a = dc.Box([0, 0, 0], [2, 4, 1.5], yaw_degrees=0)
    b = dc.Box([2, 0, 0], [2, 4, 1.5], yaw_degrees=0)            # shifted by half the length: IoU 1/3
    assert abs(M.estimateOverlap(a, b, 2, up_axis) - 1 / 3) < 1e-12 and M.estimateOverlap(a, a, 3, up_axis) == pytest.approx(1.0)
    c = dc.Box([0, 0, 0], [2, 4, 1.5], yaw_degrees=90)           # crossed: 2x2 core of two 2x4 rectangles
    assert abs(M.estimateOverlap(a, c, 2, up_axis) - 4 / 12) < 1e-12
    s, p = M.Success(), M.Precision()
    ov, ac = rng.uniform(0, 1, 50), rng.uniform(0, 3, 50)
    s(ov), p(ac)
    assert abs(s.compute() - R.success(ov)) < 1e-9 and abs(p.compute() - R.precision(ac)) < 1e-9


class _Echo(MatchingBaseModel):
    """Stand-in network: proposes the offset that re-centres the search-area's point mass (exercises the plumbing)."""

    def __init__(self, cfg):
        super().__init__(cfg)
        self.dummy = torch.nn.Parameter(torch.zeros(1))

    @property
    def device(self):
        return self.dummy.device

    def forward(self, d):
        s = d["search_points"][0]
        near = s[(s.abs() < torch.tensor([2.6, 1.4, 1.2])).all(-1)]
        c = near.mean(0) if near.shape[0] > 0 else torch.zeros(3)
        box = torch.stack([c[0] + 0.55, c[1], torch.zeros(()), torch.zeros(()), torch.ones(())])
        return {"estimation_boxes": box[None, None, :]}


def _cfg(**kw):
    c = dict(search_bb_scale=1.25, search_bb_offset=2, model_bb_scale=1.25, model_bb_offset=0, template_size=512, search_size=1024,
             degrees=True, use_z=True, limit_box=False, IoU_space=3, up_axis=[0, 0, 1], reference_BB="previous_result",
             shape_aggregation="firstandprevious")
    c.update(kw)
    return EasyDict(c)


def test_frame_loop_follows_a_synthetic_tracklet():
    seq = synthetic_sequence(n_frames=6, n_points=6000, seed=5)
    m = _Echo(_cfg())
    ious, dists, boxes = m.evaluate_one_sequence(seq)
    assert len(ious) == len(dists) == len(boxes) == 6 and ious[0] == pytest.approx(1.0) and dists[0] == 0
    assert min(ious) > 0.3 and max(dists) < 1.0                   # the stand-in keeps the object inside the window
    data, ref = m.build_input_dict(seq, 3, boxes[:3])
    assert data["template_points"].shape == (1, 512, 3) and data["search_points"].shape == (1, 1024, 3)


def test_device_tracker_matches_the_host_loop_on_cpu():
    seq = synthetic_sequence(n_frames=5, n_points=6000, seed=7)
    m = _Echo(_cfg())
    _, _, host = m.evaluate_one_sequence(seq)
    trk = DeviceTracker(m, max_points=6000, use_graph=False)
    pts = [torch.tensor(f["pc"].points.T.copy()) for f in seq]
    trk.reset(pts[0], seq[0]["3d_bbox"].to_tensor())
    for i in range(1, 5):
        b = trk.step(pts[i])
        # different random subsets of the same crops -> the stand-in's centroid estimate agrees to a few centimetres
        assert np.abs(b.center.numpy() - host[i].center).max() < 0.15
        assert np.abs(b.rot.numpy() - host[i].rotation_matrix).max() < 1e-6


def test_motion_input_matches_reference_restatement():
    """MotionBaseModel.build_input_dict (M2-Track) against the numpy restatement, first and later frames."""
    import os
    from open3dsot_b200.config import load_config
    from open3dsot_b200.models.base_model import MotionBaseModel
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = load_config(os.path.join(root, "cfgs", "M2_track_kitti.yaml"), {})
    seq = synthetic_sequence(n_frames=4, n_points=5000, seed=9)

    class Host(MotionBaseModel):
        device = torch.device("cpu")
    m = Host(cfg)
    for frame_id in (1, 3):
        ref = seq[frame_id - 1]["3d_bbox"]
        data, _ = m.build_input_dict(seq, frame_id, [ref])
        want = R.motion_build_input(seq[frame_id - 1]["pc"].points.astype(np.float64), seq[frame_id]["pc"].points.astype(np.float64),
                                    R.Box(ref.center, ref.wlh, ref.rotation_matrix), cfg, frame_id)
        for k in want:
            assert data[k].shape == want[k].shape and np.abs(data[k].numpy() - want[k]).max() < 2e-5, k
        vals = np.unique(data["points"][0, : cfg.point_sample_size, 4].numpy().astype(np.float64).round(3))
        assert set(vals.tolist()) <= ({0.0, 1.0} if frame_id == 1 else {0.2, 0.8})


def test_device_tracker_motion_mode_inputs():
    """M2-Track style two-frame input on the tracker's static buffers: channel layout, prior-targetness values (1/0 on the
    first tracked frame, 0.8/0.2 afterwards) and the BoxCloud halves."""
    from open3dsot_b200.models.base_model import MotionBaseModel
    cfg = EasyDict(point_sample_size=256, bb_scale=1.25, bb_offset=2, box_aware=True, degrees=False, use_z=True, limit_box=False,
                   IoU_space=3, up_axis=[0, 0, 1])

    class Stub(MotionBaseModel):
        def __init__(self, c):
            super().__init__(c)
            self.dummy = torch.nn.Parameter(torch.zeros(1))
            self.seen = []

        @property
        def device(self):
            return self.dummy.device

        def forward(self, d):
            self.seen.append({k: v.clone() for k, v in d.items()})
            return {"estimation_boxes": torch.tensor([[0.55, 0.0, 0.0, 0.02]])}
    m = Stub(cfg)
    seq = synthetic_sequence(n_frames=4, n_points=4000, seed=17)
    trk = DeviceTracker(m, max_points=4000, use_graph=False)
    assert trk.motion
    pts = [torch.tensor(f["pc"].points.T.copy()) for f in seq]
    trk.reset(pts[0], seq[0]["3d_bbox"].to_tensor())
    for i in range(1, 4):
        b = trk.step(pts[i])
    n = cfg.point_sample_size
    for i, d in enumerate(m.seen):
        p = d["points"][0]
        assert p.shape == (2 * n, 5) and d["candidate_bc"].shape == (1, 2 * n, 9)
        assert float(p[:n, 3].abs().sum()) == 0 and torch.allclose(p[n:, 3], torch.full((n,), 0.1)) and torch.allclose(p[n:, 4], torch.full((n,), 0.5))
        vals = set(p[:n, 4].double().round(decimals=3).unique().tolist())
        assert vals <= ({0.0, 1.0} if i == 0 else {0.2, 0.8}) and len(vals) == 2
        assert float(d["candidate_bc"][0, n:].abs().sum()) == 0 and float(d["candidate_bc"][0, :n].min()) > 0
    _, _, host = m.evaluate_one_sequence(seq)                    # the reference-shaped loop with the same constant offsets
    assert np.abs(b.center.numpy() - host[3].center).max() < 1e-4


def test_points_utils_names_match_the_restatement():
    """datasets/points_utils.py under the reference's names and signatures (host containers in, host containers out)."""
    from open3dsot_b200.datasets import points_utils as PU
    rng, ob, tb, pts = _pair(3)
    pc, hb = dc.PointCloud(pts.copy()), dc.Box(ob.center, ob.wlh, ob.rot)
    want, wbox = R.crop_and_center_pc(pts, ob, offset=0.2, scale=1.25)
    got, gbox = PU.cropAndCenterPC(pc, hb, offset=0.2, scale=1.25)
    assert got.points.shape == want.shape and np.abs(got.points - want).max() < 1e-12 and np.abs(gbox.center).max() < 1e-12
    assert np.abs(PU.generate_subwindow(pc, hb, 1.25, 2).points - R.generate_subwindow(pts, ob, 1.25, 2)).max() < 1e-12
    w2, _ = R.crop_pc_axis_aligned(pts, ob, offset=0.5, scale=1.1)
    g2, mask = PU.crop_pc_axis_aligned(pc, hb, offset=0.5, scale=1.1, return_mask=True)
    assert np.array_equal(g2.points, w2) and int(mask.sum()) == w2.shape[1]
    ob2 = R.get_offset_bb(ob, [0.3, 0.1, 0.0, 7.0], use_z=True, limit_box=False)
    hb2 = PU.getOffsetBB(hb, [0.3, 0.1, 0.0, 7.0], use_z=True, limit_box=False)
    assert np.abs(hb2.center - ob2.center).max() < 1e-12 and np.abs(hb2.rotation_matrix - ob2.rot).max() < 1e-12
    wm, wb = R.get_model([pts, pts + 0.01], [ob, ob2], offset=0, scale=1.25)
    gm, gb = PU.getModel([pc, dc.PointCloud(pts + 0.01)], [hb, hb2], offset=0, scale=1.25)
    assert gm.points.shape == wm.shape and np.abs(gm.points - wm).max() < 1e-12
    assert np.abs(PU.get_point_to_box_distance(pts.T[:50], hb) - R.get_point_to_box_distance(pts.T[:50], ob)).max() < 1e-12
    tbx = PU.transform_box(hb2, hb)
    wtb = R.transform_box(ob2, ob)
    assert np.abs(tbx.center - wtb.center).max() < 1e-12 and np.abs(tbx.rotation_matrix - wtb.rot).max() < 1e-12
    assert np.array_equal(PU.get_in_box_mask(pc, hb), R.get_in_box_mask(pts, ob))
    assert np.abs(PU.transform_pc(pc, hb).points - ob.rot.T @ (pts - ob.center[:, None])).max() < 1e-12
    p, i = PU.regularize_pc(pts.T, 128, seed=1)
    wp, wi = R.regularize_pc(pts.T, 128, seed=1)
    assert np.array_equal(i, wi) and np.array_equal(p, wp)


def test_evaluate_accumulates_success_and_precision():
    from open3dsot_b200.tracking.evaluate import evaluate
    seqs = [synthetic_sequence(n_frames=4, n_points=4000, seed=60 + i) for i in range(2)]
    m = _Echo(_cfg())
    out = evaluate(m, seqs)
    ious, dists = [], []
    for s in seqs:
        i, d, _ = m.evaluate_one_sequence(s)
        ious += i
        dists += d
    assert out["frames"] == 8 and len(out["results"]) == 2
    assert abs(out["success"] - R.success(ious)) < 1e-9 and abs(out["precision"] - R.precision(dists)) < 1e-9
    assert 30 < out["success"] <= 100 and 30 < out["precision"] <= 100
