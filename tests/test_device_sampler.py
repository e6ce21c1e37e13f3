"""On-device construction of siamese training batches against the numpy restatement of
datasets/sampler.py:16-79 (oracle/tracking_ref.py; parity unpinned, see its header), with the random draws shared."""
import os

import numpy as np
import pytest
import torch

from oracle import tracking_ref as R
from open3dsot_b200.config import load_config
from open3dsot_b200.datasets.device_sampler import DeviceSiameseSampler, DeviceTracklets, siamese_batch
from open3dsot_b200.datasets.synthetic import synthetic_sequence

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rows_in(rows, pool, tol=2e-5):
    """every row of `rows` (k, 3) occurs in `pool` (n, 3)"""
    d = np.abs(rows[:, None, :] - pool[None, :, :]).max(-1)
    return d.min(1).max() < tol, d.argmin(1)


@pytest.mark.parametrize("cfg_name,augment", [("BAT_Car.yaml", False), ("P2B_Car.yaml", False), ("BAT_Car.yaml", True)])
def test_batch_matches_reference_restatement(cfg_name, augment):
    cfg = load_config(os.path.join(ROOT, "cfgs", cfg_name), {"use_augmentation": augment})
    tracklets = [synthetic_sequence(n_frames=4, n_points=3000, seed=21 + i, n_object=400) for i in range(2)]
    data = DeviceTracklets(tracklets, "cpu")
    frames = [f for t in tracklets for f in t]
    B = 8
    g = torch.Generator().manual_seed(5)
    frame_ids = torch.tensor([0, 1, 2, 3, 4, 5, 6, 7])
    cand = torch.tensor([0, 1, 2, 3, 0, 1, 2, 3])
    draws = {"template_offset": torch.rand(B, 3, generator=g) * 0.6 - 0.3, "search_offset": torch.randn(B, 3, generator=g) * 0.6,
             "limit_rand_t": torch.rand(B, 2, generator=g) * 2 - 1, "limit_rand_s": torch.rand(B, 2, generator=g) * 2 - 1}
    if augment:
        u = torch.rand(B, 6, generator=g)
        draws["aug_search"] = {"trans": u[:, :3] * 0.6 - 0.3, "rot": u[:, 3] * 20 - 10, "flip_x": u[:, 4] < 0.5, "flip_y": u[:, 5] < 0.5}
    batch, valid = siamese_batch(data, cfg, frame_ids, cand, draws=draws, generator=g)
    assert bool(valid.all())
    deg = 5.0 if cfg.degrees else np.deg2rad(5.0)
    def fr(f):
        bb = f["3d_bbox"]
        return f["pc"].points.astype(np.float64), R.Box(bb.center, bb.wlh, bb.rotation_matrix)

    for b in range(B):
        k = int(frame_ids[b])
        first, prev = frames[int(data.first[k])], frames[int(data.prev[k])]
        off_s = draws["search_offset"][b].double().numpy() * np.sqrt([1.0, 1.0, deg])     # N(0, diag(1, 1, 5 deg))
        want = R.siamese_processing(fr(first), fr(prev), fr(frames[k]), int(cand[b]), cfg,
                                    draws["template_offset"][b].double().numpy(), off_s,
                                    limit_rand_t=draws["limit_rand_t"][b].tolist(), limit_rand_s=draws["limit_rand_s"][b].tolist(),
                                    aug_search=None if not augment else (draws["aug_search"]["trans"][b].double().numpy(),
                                                                         float(draws["aug_search"]["rot"][b]), bool(draws["aug_search"]["flip_x"][b]),
                                                                         bool(draws["aug_search"]["flip_y"][b])))
        assert int(batch["_n_template"][b]) == want["n_template"] and int(batch["_n_search"][b]) == want["n_search"]
        assert np.abs(batch["box_label"][b].numpy() - want["box_label"]).max() < 1e-4
        assert np.abs(batch["bbox_size"][b].numpy() - want["bbox_size"]).max() < 1e-6
        # the resampled clouds are subsets of the reference's crops (a different random subset), with the reference's labels
        ok, _ = _rows_in(batch["template_points"][b].numpy(), want["_model_pc"].T)
        assert ok
        ok, where = _rows_in(batch["search_points"][b].numpy(), want["_crop"].T)
        assert ok
        assert np.array_equal(batch["seg_label"][b].numpy() > 0.5, want["_seg"][where].astype(bool))
        assert 0 < batch["seg_label"][b].sum() < cfg.search_size                # object and background both present
        if cfg.get("box_aware", False):
            tp, sp = batch["template_points"][b].double().numpy(), batch["search_points"][b].double().numpy()
            assert np.abs(batch["points2cc_dist_t"][b].numpy() - R.get_point_to_box_distance(tp, want["_model_box"])).max() < 1e-4
            assert np.abs(batch["points2cc_dist_s"][b].numpy() - R.get_point_to_box_distance(sp, want["_search_box"])).max() < 1e-4


def test_sampler_returns_reference_batch_schema():
    cfg = load_config(os.path.join(ROOT, "cfgs", "BAT_Car.yaml"), {"batch_size": 6})
    tracklets = [synthetic_sequence(n_frames=5, n_points=3000, seed=3 + i, n_object=400) for i in range(3)]
    smp = DeviceSiameseSampler(tracklets, cfg, "cpu", seed=1)
    batch, valid = smp.next_batch()
    assert bool(valid.all())
    want = {"template_points": (6, cfg.template_size, 3), "search_points": (6, cfg.search_size, 3), "box_label": (6, 4),
            "bbox_size": (6, 3), "seg_label": (6, cfg.search_size), "points2cc_dist_t": (6, cfg.template_size, 9),
            "points2cc_dist_s": (6, cfg.search_size, 9)}
    assert {k: tuple(v.shape) for k, v in batch.items()} == want
    assert all(v.dtype == torch.float32 for v in batch.values())
    b2, _ = smp.next_batch()
    assert not torch.equal(b2["search_points"], batch["search_points"])      # a new draw every call


def test_motion_batch_matches_reference_restatement():
    from open3dsot_b200.datasets.device_sampler import DeviceMotionSampler, motion_batch
    cfg = load_config(os.path.join(ROOT, "cfgs", "M2_track_kitti.yaml"), {})
    tracklets = [synthetic_sequence(n_frames=4, n_points=3000, seed=31 + i, n_object=400) for i in range(2)]
    data = DeviceTracklets(tracklets, "cpu")
    frames = [f for t in tracklets for f in t]
    B = 8
    g = torch.Generator().manual_seed(6)
    frame_ids, cand = torch.arange(B), torch.tensor([0, 1, 2, 3, 0, 1, 2, 3])
    def aug():
        u = torch.rand(B, 6, generator=g)
        return {"trans": u[:, :3] * 0.6 - 0.3, "rot": u[:, 3] * 20 - 10, "flip_x": u[:, 4] < 0.5, "flip_y": u[:, 5] < 0.5}
    assert cfg.use_augmentation                                    # M2_track_kitti.yaml trains with the augmentation transform
    draws = {"offset": torch.rand(B, 3, generator=g) * 0.6 - 0.3, "limit_rand": torch.rand(B, 2, generator=g) * 2 - 1,
             "aug_prev": aug(), "aug_this": aug()}
    batch, valid = motion_batch(data, cfg, frame_ids, cand, draws=draws, generator=g)
    assert bool(valid.all())
    n = cfg.point_sample_size

    def fr(f):
        bb = f["3d_bbox"]
        return f["pc"].points.astype(np.float64), R.Box(bb.center, bb.wlh, bb.rotation_matrix)

    for b in range(B):
        k = int(frame_ids[b])
        a = lambda d: (d["trans"][b].double().numpy(), float(d["rot"][b]), bool(d["flip_x"][b]), bool(d["flip_y"][b]))
        want = R.motion_processing(fr(frames[int(data.prev[k])]), fr(frames[k]), int(cand[b]), cfg, draws["offset"][b].double().numpy(),
                                   limit_rand=draws["limit_rand"][b].tolist(), aug_prev=a(draws["aug_prev"]), aug_this=a(draws["aug_this"]))
        assert (int(batch["_n_prev"][b]), int(batch["_n_this"][b]), int(batch["_n_target"][b])) == (want["n_prev"], want["n_this"], want["n_target"])
        for key in ("box_label", "box_label_prev", "motion_label", "bbox_size"):
            assert np.abs(batch[key][b].numpy() - want[key]).max() < 2e-4, key
        assert int(batch["motion_state_label"][b]) == want["motion_state_label"]
        pts = batch["points"][b].numpy()
        ok_p, _ = _rows_in(pts[:n, :3], want["_prev_crop"].T)
        ok_t, _ = _rows_in(pts[n:, :3], want["_this_crop"].T)
        assert ok_p and ok_t
        assert np.all(pts[:n, 3] == 0) and np.allclose(pts[n:, 3], 0.1) and np.allclose(pts[n:, 4], 0.5)
        this_b, prev_b, ref_b = want["_boxes"]
        # labels recomputed by the restatement on the device's own (different random) subset
        seg = np.hstack([R._in_box_inclusive(pts[:n, :3].astype(np.float64), prev_b, 1.25), R._in_box_inclusive(pts[n:, :3].astype(np.float64), this_b, 1.25)])
        flips = int((batch["seg_label"][b].numpy().astype(bool) != seg).sum())
        assert flips <= 2                                                      # float32 vs float64 on the box faces
        m = R._in_box_inclusive(pts[:n, :3].astype(np.float64), ref_b, 1.25).astype(float)
        if int(cand[b]) != 0:
            m = np.where(m == 0, 0.2, 0.8)
        assert int((np.abs(pts[:n, 4] - m) > 1e-6).sum()) <= 2
        assert np.abs(batch["prev_bc"][b].numpy() - R.get_point_to_box_distance(pts[:n, :3].astype(np.float64), prev_b)).max() < 1e-4
        assert np.abs(batch["this_bc"][b].numpy() - R.get_point_to_box_distance(pts[n:, :3].astype(np.float64), this_b)).max() < 1e-4
        assert np.abs(batch["candidate_bc"][b, :n].numpy() - R.get_point_to_box_distance(pts[:n, :3].astype(np.float64), ref_b)).max() < 1e-4
        assert float(batch["candidate_bc"][b, n:].abs().sum()) == 0

    smp = DeviceMotionSampler(tracklets, load_config(os.path.join(ROOT, "cfgs", "M2_track_kitti.yaml"), {"batch_size": 4}), "cpu", seed=2)
    out, valid = smp.next_batch()
    from open3dsot_b200.datasets.synthetic import synthetic_motion_batch
    ref = synthetic_motion_batch(4, cfg.point_sample_size)
    assert {k: (tuple(v.shape), v.dtype) for k, v in out.items()} == {k: (tuple(v.shape), v.dtype) for k, v in ref.items()}


def test_augmentation_matches_reference_restatement():
    from open3dsot_b200.datasets.device_sampler import apply_augmentation
    from open3dsot_b200.tracking import boxes as bx
    seq = synthetic_sequence(n_frames=4, n_points=3000, seed=51, n_object=500)
    B = 4
    pts = torch.stack([torch.tensor(f["pc"].points.T.astype(np.float64)) for f in seq])
    boxes = bx.Box(*(torch.stack([getattr(f["3d_bbox"].to_tensor(dtype=torch.float64), k) for f in seq]) for k in ("center", "wlh", "rot")))
    g = torch.Generator().manual_seed(2)
    trans = torch.rand(B, 3, generator=g, dtype=torch.float64) * 0.6 - 0.3
    rot = torch.rand(B, generator=g, dtype=torch.float64) * 20 - 10
    fx, fy = torch.tensor([True, False, True, False]), torch.tensor([True, True, False, False])
    got_p, got_b = apply_augmentation(pts, boxes, trans, rot, fx, fy)
    for b in range(B):
        ob = R.Box(seq[b]["3d_bbox"].center, seq[b]["3d_bbox"].wlh, seq[b]["3d_bbox"].rotation_matrix)
        want_p, want_b = R.apply_augmentation(seq[b]["pc"].points.astype(np.float64), ob, trans[b].numpy(), float(rot[b]), bool(fx[b]), bool(fy[b]))
        assert np.abs(got_p[b].numpy().T - want_p).max() < 1e-9
        assert np.abs(got_b.center[b].numpy() - want_b.center).max() < 1e-9 and np.abs(got_b.rot[b].numpy() - want_b.rot).max() < 1e-9
        moved = np.abs(want_p - seq[b]["pc"].points).max(0) > 1e-9
        assert 300 < int(moved.sum()) < 900                         # the object's points moved, the background did not


def test_ragged_scans_and_single_candidate():
    """Scans of different lengths are padded (padding rows never survive a crop); num_candidates = 1 always offsets the
    search box (sampler.py:50-53)."""
    cfg = load_config(os.path.join(ROOT, "cfgs", "BAT_Car.yaml"), {"num_candidates": 1})
    tracklets = [synthetic_sequence(n_frames=3, n_points=n, seed=70 + i, n_object=300) for i, n in enumerate((2000, 3500))]
    data = DeviceTracklets(tracklets, "cpu")
    assert data.scans.shape == (6, 3500, 3) and data.count.tolist() == [2000] * 3 + [3500] * 3
    g = torch.Generator().manual_seed(0)
    frame_ids, cand = torch.arange(6), torch.zeros(6, dtype=torch.long)
    batch, valid = siamese_batch(data, cfg, frame_ids, cand, generator=g)
    assert bool(valid.all())
    assert float(batch["box_label"][:, 3].abs().max()) > 0            # candidate 0 of a 1-candidate sampler is still offset
    # no padded (all-zero) row of the short scans made it into a search cloud: the origin lies outside every sub-window here
    assert float(batch["search_points"][:3].abs().sum(-1).min()) > 0
    assert int(batch["_n_search"][:3].max()) <= 2000 and int(batch["_n_template"][:3].max()) <= 4000   # survivors come from real rows only
