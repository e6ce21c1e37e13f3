"""KITTI tracking reader with the reference's class and method names (datasets/kitti.py:16-205, base_dataset.py): scene
lists per split, tracklets grouped by `track_id` and ordered by frame, per-frame `{"pc": PointCloud, "3d_bbox": Box,
"meta": anno}` with the box converted from the camera-frame label to the velodyne frame through `Tr_velo_cam`.

Differences in form, not in result: label files are parsed with plain string splitting (no pandas dependency in the data
path), orientations are rotation matrices (`Quaternion(axis=[0,0,-1], radians=a)` is a rotation by -a about +z), and
`tracklets()` hands the whole split to `DeviceTracklets` / `DeviceSiameseSampler` so that batches are then built on the
device.  The directory layout is the reference's: <path>/{velodyne/<scene>/<frame:06>.bin, label_02/<scene>.txt,
calib/<scene>.txt}."""
import os
import pickle
from collections import defaultdict

import numpy as np

from .data_classes import Box, PointCloud

_COLUMNS = ("frame", "track_id", "type", "truncated", "occluded", "alpha", "bbox_left", "bbox_top", "bbox_right", "bbox_bottom",
            "height", "width", "length", "x", "y", "z", "rotation_y")
_KNOWN = ('Car', 'Van', 'Truck', 'Pedestrian', 'Person_sitting', 'Cyclist', 'Tram', 'Misc')


def _rotz(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def _rot(axis, a):
    axis = np.asarray(axis, dtype=np.float64) / np.linalg.norm(axis)
    k = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(a) * k + (1 - np.cos(a)) * (k @ k)


class BaseDataset:
    def __init__(self, path, split, category_name="Car", **kwargs):
        self.path, self.split, self.category_name = path, split, category_name
        self.preloading = kwargs.get('preloading', False)


class kittiDataset(BaseDataset):
    def __init__(self, path, split, category_name="Car", **kwargs):
        super().__init__(path, split, category_name, **kwargs)
        self.KITTI_Folder = path
        self.KITTI_velo = os.path.join(path, "velodyne")
        self.KITTI_label = os.path.join(path, "label_02")
        self.KITTI_calib = os.path.join(path, "calib")
        self.scene_list = self._build_scene_list(split)
        self.velos = defaultdict(dict)
        self.calibs = {}
        self.coordinate_mode = kwargs.get('coordinate_mode', 'velodyne')
        self.preload_offset = kwargs.get('preload_offset', -1)
        self.tracklet_anno_list, self.tracklet_len_list = self._build_tracklet_anno()
        if self.preloading:
            self.training_samples = self._load_data()

    @staticmethod
    def _build_scene_list(split):
        """kitti.py:33-55: scenes 0-16 train, 17-18 validation, 19-20 test ('tiny' variants: 0 / 18 / 19), else all 21."""
        s = split.upper()
        tiny = "TINY" in s
        if "TRAIN" in s:
            names = [0] if tiny else range(0, 17)
        elif "VALID" in s:
            names = [18] if tiny else range(17, 19)
        elif "TEST" in s:
            names = [19] if tiny else range(19, 21)
        else:
            names = range(21)
        return ['%04d' % n for n in names]

    def _load_data(self):
        path = os.path.join(self.KITTI_Folder,
                            f"preload_kitti_{self.category_name}_{self.split}_{self.coordinate_mode}_{self.preload_offset}.dat")
        if os.path.isfile(path):
            with open(path, 'rb') as f:
                return pickle.load(f)
        samples = [[self._get_frame_from_anno(a) for a in annos] for annos in self.tracklet_anno_list]
        with open(path, 'wb') as f:
            pickle.dump(samples, f)
        return samples

    def get_num_scenes(self):
        return len(self.scene_list)

    def get_num_tracklets(self):
        return len(self.tracklet_anno_list)

    def get_num_frames_total(self):
        return sum(self.tracklet_len_list)

    def get_num_frames_tracklet(self, tracklet_id):
        return self.tracklet_len_list[tracklet_id]

    def _wanted(self, kind):
        c = self.category_name
        if c in _KNOWN:
            return kind == c
        if c == 'All':
            return kind in ('Car', 'Van', 'Pedestrian', 'Cyclist')
        return kind != 'DontCare'

    def _build_tracklet_anno(self):
        """kitti.py:96-133: one list of annotations per (scene, track_id), in order of first appearance, sorted by frame."""
        tracklets, lengths = [], []
        for scene in self.scene_list:
            label_file = os.path.join(self.KITTI_label, scene + ".txt")
            if not os.path.isfile(label_file):
                continue
            per_track = {}
            with open(label_file) as f:
                for line in f:
                    v = line.split()
                    if len(v) < len(_COLUMNS) or not self._wanted(v[2]):
                        continue
                    anno = {"scene": scene}
                    for name, raw in zip(_COLUMNS, v):
                        anno[name] = raw if name == "type" else (int(raw) if name in ("frame", "track_id", "truncated", "occluded")
                                                                 else float(raw))
                    per_track.setdefault(anno["track_id"], []).append(anno)
            for annos in per_track.values():
                annos.sort(key=lambda a: a["frame"])
                tracklets.append(annos)
                lengths.append(len(annos))
        return tracklets, lengths

    def get_frames(self, seq_id, frame_ids):
        if self.preloading:
            return [self.training_samples[seq_id][f] for f in frame_ids]
        annos = self.tracklet_anno_list[seq_id]
        return [self._get_frame_from_anno(annos[f]) for f in frame_ids]

    def tracklets(self):
        """Every tracklet of the split as a list of frames (the input of DeviceTracklets / DeviceSiameseSampler)."""
        return [self.get_frames(i, range(n)) for i, n in enumerate(self.tracklet_len_list)]

    def _get_frame_from_anno(self, anno):
        """kitti.py:144-188."""
        scene_id, frame_id = anno['scene'], anno['frame']
        if scene_id not in self.calibs:
            self.calibs[scene_id] = self._read_calib_file(os.path.join(self.KITTI_calib, scene_id + ".txt"))
        velo_to_cam = np.vstack((self.calibs[scene_id]["Tr_velo_cam"], np.array([0, 0, 0, 1])))
        size = [anno["width"], anno["length"], anno["height"]]
        if self.coordinate_mode == 'velodyne':
            center_cam = np.array([anno["x"], anno["y"] - anno["height"] / 2, anno["z"], 1])
            center = (np.linalg.inv(velo_to_cam) @ center_cam)[:3]
            rot = _rotz(-anno["rotation_y"]) @ _rotz(-np.pi / 2)
        else:
            center = [anno["x"], anno["y"] - anno["height"] / 2, anno["z"]]
            rot = _rot([0, 1, 0], anno["rotation_y"]) @ _rot([1, 0, 0], np.pi / 2)
        bb = Box(center, size, rot)
        try:
            if frame_id not in self.velos[scene_id]:
                path = os.path.join(self.KITTI_velo, scene_id, '{:06}.bin'.format(frame_id))
                pts = np.fromfile(path, dtype=np.float32).reshape(-1, 4).T
                pc = PointCloud(pts)
                if self.coordinate_mode == "camera":
                    pc.points = (velo_to_cam @ np.vstack((pc.points[:3], np.ones(pc.points.shape[1]))))[:3]
                self.velos[scene_id][frame_id] = pc
            pc = self.velos[scene_id][frame_id]
            if self.preload_offset > 0:                       # crop_pc_axis_aligned(pc, bb, offset=preload_offset)
                c = bb.corners()
                lo, hi = c.min(1) - self.preload_offset, c.max(1) + self.preload_offset
                keep = ((pc.points > lo[:, None]) & (pc.points < hi[:, None])).all(0)
                pc = PointCloud(pc.points[:, keep])
        except (OSError, ValueError):
            pc = PointCloud(np.array([[0, 0, 0]], dtype=np.float32).T)
        return {"pc": pc, "3d_bbox": bb, 'meta': anno}

    @staticmethod
    def _read_calib_file(filepath):
        """kitti.py:190-205: every line that parses as twelve floats becomes a 3x4 matrix keyed by its first token."""
        data = {}
        with open(filepath) as f:
            for line in f:
                v = line.split()
                try:
                    data[v[0]] = np.array([float(x) for x in v[1:]]).reshape(3, 4)
                except (ValueError, IndexError):
                    pass
        return data
