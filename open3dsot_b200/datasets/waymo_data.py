"""Waymo reader with the reference's class and method names (datasets/waymo_data.py:22-208).  Like the reference it starts from
the per-category tracklet index `sot_infos_<category>_<split>.pkl` ({tracklet name: [{"PC": path of the frame's lidar pickle,
"Box": [cx, cy, cz, l, w, h, vx, vy, heading], "Class": name}, ...]}) and the converter's per-frame pickles
(`lidar/...pkl` with `lidars.points_xyz`, `annos/...pkl` with the 4x4 `veh_to_global`); producing those from the raw
tfrecords is the reference's offline conversion script (datasets/generate_waymo_sot.py) and is out of scope here — a missing
index raises FileNotFoundError instead of starting a conversion.

Per frame (waymo_data.py:121-168): points vehicle -> global with `veh_to_global`; the box is built in the vehicle frame with
width / length swapped into the (w, l, h) order and a rotation of -heading about z (Waymo measures heading clockwise from +x in
this convention), then rotated and translated into the global frame."""
import os
import pickle

import numpy as np

from .data_classes import Box, PointCloud
from .kitti import BaseDataset, _rotz


class WaymoDataset(BaseDataset):
    def __init__(self, path, split, category_name="VEHICLE", **kwargs):
        super().__init__(path, split, category_name, **kwargs)
        self.Waymo_Folder = path
        self.split = 'val' if split.lower() == 'test' else split.lower()
        self.category_name = category_name.lower()
        assert self.split in ('train', 'val') and self.category_name in ('vehicle', 'pedestrian', 'cyclist')
        self.tiny = kwargs.get('tiny', False)
        self.tracklet_anno_list, self.tracklet_len_list = self._build_tracklet_anno()
        if self.tiny:
            self.tracklet_anno_list, self.tracklet_len_list = self.tracklet_anno_list[:100], self.tracklet_len_list[:100]
        self.preload_offset = kwargs.get('preload_offset', 10)
        if self.preloading:
            self.training_samples = self._load_data()

    def _build_tracklet_anno(self):
        index = os.path.join(self.Waymo_Folder, f"sot_infos_{self.category_name}_{self.split}.pkl")
        if not os.path.exists(index):
            raise FileNotFoundError(f"{index} not found: run the reference's Waymo conversion (datasets/generate_waymo_sot.py) first")
        with open(index, 'rb') as f:
            infos = pickle.load(f)
        annos = [infos[k] for k in infos.keys()]
        return annos, [len(a) for a in annos]

    def _load_data(self):
        tag = f"{self.split}_{self.category_name}_{self.preload_offset}" + ("_tiny" if self.tiny else "")
        path = os.path.join(self.Waymo_Folder, f"preload_{tag}.dat")
        if os.path.isfile(path):
            with open(path, 'rb') as f:
                return pickle.load(f)
        samples = [[self._get_frame_from_anno(a) for a in annos] for annos in self.tracklet_anno_list]
        with open(path, 'wb') as f:
            pickle.dump(samples, f)
        return samples

    def get_num_tracklets(self):
        return len(self.tracklet_anno_list)

    def get_num_frames_total(self):
        return sum(self.tracklet_len_list)

    def get_num_frames_tracklet(self, tracklet_id):
        return self.tracklet_len_list[tracklet_id]

    def get_frames(self, seq_id, frame_ids):
        if self.preloading:
            return [self.training_samples[seq_id][f] for f in frame_ids]
        annos = self.tracklet_anno_list[seq_id]
        return [self._get_frame_from_anno(annos[f]) for f in frame_ids]

    def tracklets(self):
        return [self.get_frames(i, range(n)) for i, n in enumerate(self.tracklet_len_list)]

    def _get_frame_from_anno(self, anno, track_id=None):
        lidar_path = anno['PC']
        gt = np.array(anno['Box'], dtype=np.float64)
        with open(lidar_path, 'rb') as f:
            pts = np.asarray(pickle.load(f)['lidars']['points_xyz'], dtype=np.float64).T            # (3, N), vehicle frame
        with open(lidar_path.replace('lidar', 'annos'), 'rb') as f:
            pose = np.reshape(pickle.load(f)['veh_to_global'], [4, 4]).astype(np.float64)
        R, t = pose[:3, :3], pose[:3, 3]                                                              # veh_pos_to_transform (:170-208)
        pts = R @ pts + t[:, None]
        size = [gt[4], gt[3], gt[5]]                                                                  # (l, w, h) -> (w, l, h)
        bb = Box(R @ gt[0:3] + t, size, R @ _rotz(-gt[-1]))
        pc = PointCloud(pts.astype(np.float32))
        if self.preload_offset > 0:
            c = bb.corners()
            lo, hi = c.min(1) - self.preload_offset, c.max(1) + self.preload_offset
            pc = PointCloud(pc.points[:, ((pc.points > lo[:, None]) & (pc.points < hi[:, None])).all(0)])
        return {"pc": pc, "3d_bbox": bb, 'meta': anno}
