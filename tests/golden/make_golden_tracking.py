"""Golden vectors for the host geometry around the network, produced by running the REFERENCE's own,
unmodified `datasets/points_utils.py`, `datasets/data_classes.py` and `datasets/sampler.py` (siamese_processing, :16-79) from
/root/reference on synthetic tracklets.  The reference's absent dependencies (pyquaternion, nuscenes-devkit's points_in_box,
easydict, pomegranate) are the numpy stand-ins of tests/golden/_ref_shims.py; its random draws (np.random.uniform in
siamese_processing / getOffsetBB, KalmanFiltering.sample, regularize_pc's index draw) are recorded and stored with the outputs,
so the oracle (oracle/tracking_ref.py) and the product (open3dsot_b200/tracking, datasets/device_sampler.py, csrc/geometry.cu)
can be replayed on exactly the same draws.

Run only inside the authoring container:   python tests/golden/make_golden_tracking.py   -> tests/golden/ref_tracking.npz"""
import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import _ref_shims  # noqa: E402
from open3dsot_b200.config import load_config  # noqa: E402
from open3dsot_b200.datasets.synthetic import synthetic_sequence  # noqa: E402


def main():
    assert os.path.isdir("/root/reference"), "golden vectors can only be generated where /root/reference exists"
    _ref_shims.install()
    from datasets import data_classes as rd, points_utils as rp, sampler as rs     # the reference's own modules
    Quaternion = _ref_shims.Quaternion
    out = {}

    def ref_frame(f):
        b = f["3d_bbox"]
        return {"pc": rd.PointCloud(f["pc"].points.astype(np.float64).copy()),
                "3d_bbox": rd.Box(np.array(b.center, dtype=np.float64), np.array(b.wlh, dtype=np.float64), Quaternion(matrix=b.rotation_matrix))}

    def put_box(key, box):
        out[key + ".center"], out[key + ".wlh"], out[key + ".rot"] = np.array(box.center), np.array(box.wlh), np.array(box.rotation_matrix)

    tracklets = [synthetic_sequence(n_frames=4, n_points=3000, seed=21 + i, n_object=400) for i in range(2)]
    frames = [f for t in tracklets for f in t]
    out["meta.tracklet_seeds"] = np.array([21, 22])
    out["meta.n_frames"], out["meta.n_points"], out["meta.n_object"] = np.array(4), np.array(3000), np.array(400)
    rf = [ref_frame(f) for f in frames]

    # ---- points_utils, function by function (frame 1 of tracklet 0 against its own / a displaced box)
    f0, f1 = rf[0], rf[1]
    box = f1["3d_bbox"]
    for i, (off, deg, use_z, limit) in enumerate([((0.1, -0.2, 3.0), True, False, True), ((0.25, 0.1, 0.04), False, False, True),
                                                  ((0.2, -0.1, 0.3, 4.0), True, True, False), ((5.0, 3.0, 2.0), True, False, True)]):
        np.random.seed(100 + i)
        nb = rp.getOffsetBB(copy.deepcopy(box), np.array(off, dtype=np.float64), degrees=deg, use_z=use_z, limit_box=limit)
        np.random.seed(100 + i)
        out[f"offset{i}.rand"] = np.array([np.random.uniform(-1, 1), np.random.uniform(-1, 1)])   # what limit_box would draw, in order
        out[f"offset{i}.args"] = np.array(list(off) + [float(deg), float(use_z), float(limit)])
        put_box(f"offset{i}.out", nb)
    moved = rp.getOffsetBB(copy.deepcopy(box), np.array([0.3, -0.2, 4.0]), degrees=True, limit_box=False)
    put_box("box_in", box)
    put_box("box_moved", moved)
    sub = rp.generate_subwindow(f1["pc"], moved, scale=1.25, offset=2)
    out["subwindow.points"] = sub.points
    model_pc, model_box = rp.getModel([f0["pc"], f1["pc"]], [f0["3d_bbox"], moved], scale=1.25, offset=0)
    out["model.points"] = model_pc.points
    put_box("model.box", model_box)
    cpc, cbox = rp.cropAndCenterPC(f1["pc"], moved, offset=0.5, scale=1.1)
    out["cropcenter.points"] = cpc.points
    put_box("cropcenter.box", cbox)
    reg, idx = rp.regularize_pc(sub.points.T, 1024, seed=1)
    out["regularize.points"], out["regularize.idx"] = reg, idx
    reg2, idx2 = rp.regularize_pc(model_pc.points.T, 512, seed=1)
    out["regularize_up.points"], out["regularize_up.idx"] = reg2, idx2              # fewer points than requested: with replacement
    out["boxcloud"] = rp.get_point_to_box_distance(reg2, model_box)
    tb = rp.transform_box(f1["3d_bbox"], moved)
    put_box("transform_box", tb)
    out["in_box_mask"] = rp.get_in_box_mask(sub, tb)
    out["points_in_box_1.25"] = _ref_shims.points_in_box(box, f1["pc"].points, 1.25)

    # ---- siamese_processing with its random draws recorded (sampler.py:16-79)
    for cfg_name in ("BAT_Car.yaml", "P2B_Car.yaml"):
        cfg = load_config(os.path.join(ROOT, "cfgs", cfg_name))
        tag = cfg_name.split("_")[0].lower()
        for s, (k, cand) in enumerate([(1, 0), (2, 1), (3, 2), (5, 0), (6, 3), (7, 1)]):
            first = rf[0] if k < 4 else rf[4]
            prev = rf[max(k - 1, 0 if k < 4 else 4)]
            draws = {"uniform": [], "kalman": None, "idx": [], "limit": []}
            np.random.seed(1000 + 10 * s + cand)
            o_uniform, o_kalman, o_reg, o_gob = np.random.uniform, rs.KalmanFiltering.sample, rp.regularize_pc, rp.getOffsetBB

            def gob(box, offset, **kw):
                # which of limit_box's two substitutions fired, and with which uniform(-1, 1) number (points_utils.py:70-76)
                n0, off_in = len(draws["uniform"]), np.array(offset, dtype=np.float64).copy()
                r = o_gob(box, offset, **kw)
                it = iter(draws["uniform"][n0:])
                rand = [np.nan, np.nan]
                if kw.get("limit_box", True):
                    if off_in[0] > box.wlh[0]:
                        rand[0] = float(next(it)[0])
                    if off_in[1] > min(box.wlh[1], 2):
                        rand[1] = float(next(it)[0])
                draws["limit"].append(rand)
                return r

            def uniform(low=0.0, high=1.0, size=None):
                v = o_uniform(low, high, size)
                draws["uniform"].append(np.atleast_1d(np.array(v, dtype=np.float64)).copy())
                return v

            def kalman(self, n=10):
                v = o_kalman(self, n)
                draws["kalman"] = np.array(v[0], dtype=np.float64).copy()
                return v

            def reg_pc(points, sample_size, seed=None):
                p, i = o_reg(points, sample_size, seed)
                draws["idx"].append(np.array(i))
                return p, i
            np.random.uniform, rs.KalmanFiltering.sample, rp.regularize_pc, rp.getOffsetBB = uniform, kalman, reg_pc, gob
            try:
                d = rs.siamese_processing({"first_frame": copy.deepcopy(first), "template_frame": copy.deepcopy(prev),
                                           "search_frame": copy.deepcopy(rf[k]), "candidate_id": cand}, cfg)
            finally:
                np.random.uniform, rs.KalmanFiltering.sample, rp.regularize_pc, rp.getOffsetBB = o_uniform, o_kalman, o_reg, o_gob
            key = f"siamese.{tag}.{s}"
            out[key + ".frames"] = np.array([frames.index(frames[0 if k < 4 else 4]), max(k - 1, 0 if k < 4 else 4), k, cand])
            u = draws["uniform"]
            out[key + ".template_offset"] = u[0] if cand != 0 else np.zeros(3)           # the uniform(-0.3, 0.3, 3) triple (raw, before * 5 deg)
            out[key + ".limit_rand_t"], out[key + ".limit_rand_s"] = np.array(draws["limit"][0]), np.array(draws["limit"][1])
            out[key + ".search_offset"] = draws["kalman"] if draws["kalman"] is not None else np.zeros(3)
            out[key + ".idx_t"], out[key + ".idx_s"] = draws["idx"][0], draws["idx"][1]
            for kk, v in d.items():
                out[key + ".out." + kk] = np.asarray(v)
    # ---- Waymo frame construction (datasets/waymo_data.py:114-208) by the reference's own reader on converter-format pickles
    import pickle
    import tempfile
    import types
    sys.modules["datasets.generate_waymo_sot"] = types.ModuleType("datasets.generate_waymo_sot")     # tfrecord converter: not needed
    sys.modules["datasets.generate_waymo_sot"].generate_waymo_data = None
    from datasets import waymo_data as rw
    rng = np.random.default_rng(7)
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "lidar")); os.makedirs(os.path.join(tmp, "annos"))
        infos = {"seg0_obj0": []}
        for i in range(3):
            pts = rng.uniform(-20, 20, size=(500, 3)).astype(np.float32)
            ang = 0.3 + 0.1 * i
            pose = np.eye(4)
            pose[:3, :3] = _ref_shims.Quaternion(axis=[0.1, -0.05, 1.0], radians=ang).rotation_matrix
            pose[:3, 3] = [100.0 + 2 * i, -50.0 + i, 3.0]
            box = np.array([4.0 + i, 1.0, 0.2, 4.2, 1.9, 1.6, 1.0, 0.5, 0.4 + 0.05 * i], dtype=np.float32)
            lp = os.path.join(tmp, "lidar", f"seq_0_frame_{i}.pkl")
            with open(lp, "wb") as f:
                pickle.dump({"lidars": {"points_xyz": pts}, "frame_id": i, "scene_name": "seg0"}, f)
            with open(lp.replace("lidar", "annos"), "wb") as f:
                pickle.dump({"veh_to_global": pose.reshape(-1)}, f)
            infos["seg0_obj0"].append({"PC": lp, "Box": box.copy(), "Class": "VEHICLE"})
            out[f"waymo.{i}.points_xyz"], out[f"waymo.{i}.veh_to_global"], out[f"waymo.{i}.box"] = pts, pose, box
        with open(os.path.join(tmp, "sot_infos_vehicle_train.pkl"), "wb") as f:
            pickle.dump(infos, f)
        ds = rw.WaymoDataset(tmp, "train", "VEHICLE", preloading=False, preload_offset=10)
        for i, fr in enumerate(ds.get_frames(0, range(3))):
            out[f"waymo.{i}.out.points"] = fr["pc"].points
            put_box(f"waymo.{i}.out.box", fr["3d_bbox"])
    np.savez_compressed(os.path.join(HERE, "ref_tracking.npz"), **out)
    print("ref_tracking.npz", os.path.getsize(os.path.join(HERE, "ref_tracking.npz")) // 1024, "KiB,", len(out), "arrays")


if __name__ == "__main__":
    main()
