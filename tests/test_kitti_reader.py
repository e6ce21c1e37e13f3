"""KITTI tracking reader against an on-disk fixture written in KITTI's own formats (label_02 text, calib text,
velodyne float32 .bin): tracklet grouping, camera->velodyne box conversion (datasets/kitti.py:144-188), point loading,
preload crop and cache, and the hand-over to the device sampler."""
import os

import numpy as np
import torch

from open3dsot_b200.config import load_config
from open3dsot_b200.datasets import get_dataset
from open3dsot_b200.datasets.kitti import kittiDataset
from open3dsot_b200.datasets.synthetic import synthetic_sequence

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# velodyne (x fwd, y left, z up) -> camera (x right, y down, z fwd), plus a small lever arm
TR = np.array([[0.0, -1.0, 0.0, 0.02], [0.0, 0.0, -1.0, -0.07], [1.0, 0.0, 0.0, -0.27]])


def _write_scene(root, scene, tracks, extra_dontcare=True):
    os.makedirs(os.path.join(root, "label_02"), exist_ok=True)
    os.makedirs(os.path.join(root, "calib"), exist_ok=True)
    os.makedirs(os.path.join(root, "velodyne", scene), exist_ok=True)
    with open(os.path.join(root, "calib", scene + ".txt"), "w") as f:
        f.write("P2: " + " ".join(["1"] * 12) + "\n")
        f.write("Tr_velo_cam " + " ".join(f"{v:.6f}" for v in TR.reshape(-1)) + "\n")
        f.write("date: 2011-09-26\n")
    lines = []
    n_frames = max(len(t) for _, t in tracks)
    for frame in range(n_frames):
        clouds = []
        for (tid, kind), seq in ((k, t) for k, t in tracks):
            if frame >= len(seq):
                continue
            fr = seq[frame]
            b = fr["3d_bbox"]
            yaw = np.arctan2(b.rotation_matrix[1, 0], b.rotation_matrix[0, 0])
            w, l, h = b.wlh
            cam = TR @ np.append(b.center, 1.0)
            # KITTI: location = bottom centre in the camera frame (y down), rotation_y about the camera's y axis
            lines.append((frame, f"{frame} {tid} {kind} 0 0 -1.0 0 0 50 50 {h:.6f} {w:.6f} {l:.6f} "
                                 f"{cam[0]:.6f} {cam[1] + h / 2:.6f} {cam[2]:.6f} {-yaw - np.pi / 2:.8f}"))
            clouds.append(fr["pc"].points.T)
        if extra_dontcare:
            lines.append((frame, f"{frame} -1 DontCare -1 -1 -10 0 0 1 1 -1 -1 -1 -1000 -1000 -1000 -10"))
        pts = np.concatenate(clouds) if clouds else np.zeros((1, 3), np.float32)
        np.concatenate([pts, np.ones((pts.shape[0], 1), np.float32)], 1).astype(np.float32).tofile(
            os.path.join(root, "velodyne", scene, f"{frame:06d}.bin"))
    with open(os.path.join(root, "label_02", scene + ".txt"), "w") as f:
        f.write("\n".join(l for _, l in sorted(lines, key=lambda x: x[0])) + "\n")


def test_reader_round_trips_a_kitti_fixture(tmp_path):
    root = str(tmp_path)
    car0 = synthetic_sequence(n_frames=4, n_points=1500, seed=1, n_object=300)
    car1 = synthetic_sequence(n_frames=3, n_points=1500, seed=2, n_object=300)
    ped = synthetic_sequence(n_frames=2, n_points=800, seed=3, wlh=(0.7, 0.9, 1.8), n_object=100)
    _write_scene(root, "0000", [((7, "Car"), car0), ((3, "Pedestrian"), ped), ((9, "Car"), car1)])
    ds = kittiDataset(root, "train_tiny", "Car", coordinate_mode="velodyne", preloading=False, preload_offset=-1)
    assert ds.scene_list == ["0000"] and ds.get_num_tracklets() == 2 and ds.tracklet_len_list == [4, 3]
    assert ds.get_num_frames_total() == 7 and ds.get_num_frames_tracklet(1) == 3
    assert [a["track_id"] for a in ds.tracklet_anno_list[0]] == [7] * 4 and [a["frame"] for a in ds.tracklet_anno_list[0]] == [0, 1, 2, 3]
    first, last = ds.get_frames(0, [0, 3])
    for got, want in ((first, car0[0]), (last, car0[3])):
        b, w = got["3d_bbox"], want["3d_bbox"]
        assert np.abs(b.center - w.center).max() < 1e-4 and np.abs(b.wlh - w.wlh).max() < 1e-5
        assert np.abs(b.rotation_matrix - w.rotation_matrix).max() < 1e-5
    assert first["pc"].points.shape[0] == 3 and first["pc"].points.shape[1] == 1500 + 800 + 1500      # whole scan of the frame
    assert kittiDataset(root, "train_tiny", "Pedestrian").tracklet_len_list == [2]
    assert kittiDataset(root, "train_tiny", "All").get_num_tracklets() == 3
    assert kittiDataset(root, "test").get_num_tracklets() == 0                                       # scenes 19-20 absent

    crop = kittiDataset(root, "train_tiny", "Car", preloading=True, preload_offset=2.0)
    f0 = crop.get_frames(0, [0])[0]
    c = f0["3d_bbox"].corners()
    assert f0["pc"].points.shape[1] < 3800
    assert (f0["pc"].points > (c.min(1) - 2.0)[:, None]).all() and (f0["pc"].points < (c.max(1) + 2.0)[:, None]).all()
    assert os.path.isfile(os.path.join(root, "preload_kitti_Car_train_tiny_velodyne_2.0.dat"))
    again = kittiDataset(root, "train_tiny", "Car", preloading=True, preload_offset=2.0)           # served from the cache file
    assert again.get_frames(0, [0])[0]["pc"].points.shape == f0["pc"].points.shape


def test_reader_feeds_the_device_sampler(tmp_path):
    root = str(tmp_path)
    _write_scene(root, "0000", [((1, "Car"), synthetic_sequence(n_frames=5, n_points=2500, seed=4, n_object=400)),
                                ((2, "Car"), synthetic_sequence(n_frames=4, n_points=2500, seed=5, n_object=400))])
    cfg = load_config(os.path.join(ROOT, "cfgs", "BAT_Car.yaml"), {"path": root, "batch_size": 4, "preloading": False})
    smp = get_dataset(cfg, type="train_siamese", split="train_tiny", device="cpu")
    batch, valid = smp.next_batch()
    assert bool(valid.all()) and batch["template_points"].shape == (4, cfg.template_size, 3)
    assert 0 < float(batch["seg_label"].sum()) < 4 * cfg.search_size and torch.isfinite(batch["points2cc_dist_s"]).all()
    seqs = get_dataset(cfg, type="test", split="train_tiny")
    assert [len(s) for s in seqs] == [5, 4] and set(seqs[0][0]) == {"pc", "3d_bbox", "meta"}


def test_camera_coordinate_mode(tmp_path):
    """coordinate_mode='camera' (kitti.py:160-166, :176-177): label kept in the camera frame, scan moved there by Tr_velo_cam."""
    root = str(tmp_path)
    seq = synthetic_sequence(n_frames=2, n_points=500, seed=8, n_object=100)
    _write_scene(root, "0000", [((4, "Car"), seq)])
    velo = kittiDataset(root, "train_tiny", "Car", coordinate_mode="velodyne").get_frames(0, [1])[0]
    cam = kittiDataset(root, "train_tiny", "Car", coordinate_mode="camera").get_frames(0, [1])[0]
    T = np.vstack((TR, [0, 0, 0, 1]))
    want_pts = (T @ np.vstack((velo["pc"].points, np.ones(velo["pc"].points.shape[1]))))[:3]
    assert np.abs(cam["pc"].points - want_pts).max() < 1e-5
    a = cam["meta"]
    assert np.allclose(cam["3d_bbox"].center, [a["x"], a["y"] - a["height"] / 2, a["z"]])
    # the same physical box: camera-frame centre = Tr_velo_cam applied to the velodyne-frame centre
    assert np.abs((T @ np.append(velo["3d_bbox"].center, 1.0))[:3] - cam["3d_bbox"].center).max() < 1e-4
    r = cam["3d_bbox"].rotation_matrix
    assert np.abs(r @ r.T - np.eye(3)).max() < 1e-12 and abs(np.linalg.det(r) - 1) < 1e-12
    # its length axis (box x) in the camera frame is the velodyne heading rotated by Tr_velo_cam
    assert np.abs(TR[:, :3] @ velo["3d_bbox"].rotation_matrix[:, 0] - r[:, 0]).max() < 1e-5
