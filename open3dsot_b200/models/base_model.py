"""Model base classes.  Mirror of models/base_model.py: BaseModel (optimizer/scheduler :28-36, the tracking frame loop
:44-86), MatchingBaseModel (compute_loss :122-164, template / search-area construction :166-247).
The frame loop keeps the reference's method names and control flow but runs its geometry as tensor math on the model's
device (open3dsot_b200/tracking/boxes.py) instead of numpy + pyquaternion on the host; `tracking.DeviceTracker` is the
fixed-shape, graph-captured form of the same loop."""
import numpy as np
import torch
import torch.nn.functional as F

from ..compat import EasyDict, LightningModule
from ..datasets import data_classes
from ..tracking import boxes as bx
from ..utils.metrics import estimateAccuracy, estimateOverlap


def _points(pc, device):
    """(N, 3) float32 tensor on `device` from a PointCloud-like object (`.points` (3, N)) or an (N, 3) array / tensor."""
    pts = getattr(pc, "points", pc)
    t = torch.as_tensor(np.asarray(pts) if not torch.is_tensor(pts) else pts, dtype=torch.float32, device=device)
    return t.t().contiguous() if (t.shape[0] == 3 and t.shape[-1] != 3) or hasattr(pc, "points") else t


def _tbox(box, device):
    return box if isinstance(box, bx.Box) else data_classes.Box(box.center, box.wlh, box.rotation_matrix).to_tensor(device)


def regularize(points, sample_size, seed=None):
    """points_utils.regularize_pc (:24-40) on an (N, 3) tensor: the index draw is the reference's numpy Generator."""
    n = points.shape[0]
    if n <= 2:
        return torch.zeros(sample_size, 3, dtype=points.dtype, device=points.device), None
    if n == sample_size:
        return points, np.arange(n)
    rng = np.random if seed is None else np.random.default_rng(seed)
    idx = rng.choice(n, size=sample_size, replace=sample_size > n)
    return points[torch.as_tensor(idx, device=points.device)], idx


class BaseModel(LightningModule):
    def __init__(self, config=None, **kwargs):
        super().__init__()
        if config is None:
            config = EasyDict(kwargs)
        self.config = config

    def configure_optimizers(self):
        if self.config.optimizer.lower() == 'sgd':
            optimizer = torch.optim.SGD(self.parameters(), lr=self.config.lr, momentum=0.9,
                                        weight_decay=self.config.wd)
        else:
            optimizer = torch.optim.Adam(self.parameters(), lr=self.config.lr, weight_decay=self.config.wd,
                                         betas=(0.5, 0.999), eps=1e-06)
        scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=self.config.lr_decay_step,
                                                    gamma=self.config.lr_decay_rate)
        return {"optimizer": optimizer, "lr_scheduler": scheduler}

    def compute_loss(self, data, output):
        raise NotImplementedError

    def build_input_dict(self, sequence, frame_id, results_bbs, **kwargs):
        raise NotImplementedError

    def evaluate_one_sample(self, data_dict, ref_box):
        """:44-60: run the network, take the best proposal, move the reference box by it (getOffsetBB)."""
        with torch.no_grad():
            end_points = self(data_dict)
        est = end_points['estimation_boxes'][0]
        if est.dim() == 2:
            est = est.index_select(0, est[:, 4].argmax().reshape(1))[0, :4]    # (indexing by a 0-d tensor would sync)
        ref = _tbox(ref_box, self.device)
        new = bx.offset_box(ref, est.to(ref.center.dtype), degrees=self.config.degrees, use_z=self.config.use_z,
                            limit_box=self.config.limit_box)
        return data_classes.Box.from_tensor(new)

    def evaluate_one_sequence(self, sequence):
        """:62-86.  sequence: list of {"pc": PointCloud, "3d_bbox": Box}; returns (ious, distances, result boxes)."""
        ious, distances, results_bbs = [], [], []
        for frame_id in range(len(sequence)):
            this_bb = sequence[frame_id]["3d_bbox"]
            if frame_id == 0:
                results_bbs.append(this_bb)
            else:
                data_dict, ref_bb = self.build_input_dict(sequence, frame_id, results_bbs)
                results_bbs.append(self.evaluate_one_sample(data_dict, ref_box=ref_bb))
            ious.append(estimateOverlap(this_bb, results_bbs[-1], dim=self.config.IoU_space, up_axis=self.config.up_axis))
            distances.append(estimateAccuracy(this_bb, results_bbs[-1], dim=self.config.IoU_space,
                                              up_axis=self.config.up_axis))
        return ious, distances, results_bbs


class MatchingBaseModel(BaseModel):
    def compute_loss(self, data, output):
        """Segmentation BCE, masked smooth-L1 vote loss, objectness BCE (pos_weight 2, thresholds 0.3 / 0.6)
        and masked smooth-L1 box loss — device-aware (the reference hard-codes `.cuda()`, base_model.py:151)."""
        estimation_boxes = output['estimation_boxes']        # (B, num_proposal, 5)
        estimation_cla = output['estimation_cla']            # (B, N)
        seg_label, box_label = data['seg_label'], data['box_label']
        proposal_center, vote_xyz = output["center_xyz"], output["vote_xyz"]

        loss_seg = F.binary_cross_entropy_with_logits(estimation_cla, seg_label)

        loss_vote = F.smooth_l1_loss(vote_xyz, box_label[:, None, :3].expand_as(vote_xyz), reduction='none')
        loss_vote = (loss_vote.mean(2) * seg_label).sum() / (seg_label.sum() + 1e-06)

        dist = torch.sqrt(torch.sum((proposal_center - box_label[:, None, :3]) ** 2, dim=-1) + 1e-6)
        objectness_label = (dist < 0.3).float()
        objectness_mask = ((dist < 0.3) | (dist > 0.6)).float()
        # the reference calls BCE with its default MEAN reduction (base_model.py:151-152): the scalar mean over all
        # proposals is what gets multiplied by the mask — kept exactly (mask only enters through sum/(sum + 1e-6))
        loss_objective = F.binary_cross_entropy_with_logits(
            estimation_boxes[:, :, 4], objectness_label,
            pos_weight=torch.full((1,), 2.0, device=estimation_boxes.device))   # device-side fill: graph-capturable
        loss_objective = torch.sum(loss_objective * objectness_mask) / (torch.sum(objectness_mask) + 1e-6)

        loss_box = F.smooth_l1_loss(estimation_boxes[:, :, :4],
                                    box_label[:, None, :4].expand_as(estimation_boxes[:, :, :4]), reduction='none')
        loss_box = torch.sum(loss_box.mean(2) * objectness_label) / (objectness_label.sum() + 1e-6)
        return {"loss_objective": loss_objective, "loss_box": loss_box, "loss_seg": loss_seg, "loss_vote": loss_vote}


    # ---- tracking input construction (:166-247) ----------------------------------------------------------------
    def _crop_and_center(self, pc, box):
        pts = _points(pc, self.device)
        local, keep, canon = bx.crop_and_center(pts, _tbox(box, self.device), offset=self.config.model_bb_offset,
                                                scale=self.config.model_bb_scale)
        return local[keep], canon

    def generate_template(self, sequence, current_frame_id, results_bbs):
        mode = self.config.shape_aggregation.upper()
        first_pc, previous_pc = sequence[0]['pc'], sequence[current_frame_id - 1]['pc']
        if "FIRSTANDPREVIOUS" in mode:
            pairs = [(first_pc, results_bbs[0]), (previous_pc, results_bbs[current_frame_id - 1])]
        elif "FIRST" in mode:
            pairs = [(first_pc, results_bbs[0])]
        elif "PREVIOUS" in mode:
            pairs = [(previous_pc, results_bbs[current_frame_id - 1])]
        elif "ALL" in mode:
            pairs = [(f["pc"], b) for f, b in zip(sequence[:current_frame_id], results_bbs)]
        else:
            raise ValueError(self.config.shape_aggregation)
        parts, canon = [], None
        for pc, box in pairs:                      # getModel (:88-100): merged crops, canonical box of the last pair
            pts, canon = self._crop_and_center(pc, box)
            parts.append(pts)
        return torch.cat(parts), canon

    def generate_search_area(self, sequence, current_frame_id, results_bbs):
        ref = self.config.reference_BB.upper()
        if "PREVIOUS_RESULT" in ref:
            ref_bb = results_bbs[-1]
        elif "PREVIOUS_GT" in ref:
            ref_bb = sequence[current_frame_id - 1]["3d_bbox"]
        elif "CURRENT_GT" in ref:
            ref_bb = sequence[current_frame_id]["3d_bbox"]
        else:
            raise ValueError(self.config.reference_BB)
        pts = _points(sequence[current_frame_id]["pc"], self.device)
        local, keep = bx.subwindow(pts, _tbox(ref_bb, self.device), scale=self.config.search_bb_scale,
                                   offset=self.config.search_bb_offset)
        return local[keep], ref_bb

    def prepare_input(self, template_pc, search_pc, template_box, *args, **kwargs):
        template_points, _ = regularize(template_pc, self.config.template_size, seed=1)
        search_points, _ = regularize(search_pc, self.config.search_size, seed=1)
        return {'template_points': template_points[None], 'search_points': search_points[None]}

    def build_input_dict(self, sequence, frame_id, results_bbs, **kwargs):
        search_pc_crop, ref_bb = self.generate_search_area(sequence, frame_id, results_bbs)
        template_pc, canonical_box = self.generate_template(sequence, frame_id, results_bbs)
        return self.prepare_input(template_pc, search_pc_crop, canonical_box), ref_bb


def motion_input(prev_local, prev_keep, this_local, this_keep, wlh, size, first_frame, box_aware, draw=None):
    """MotionBaseModel.build_input_dict (:255-303) from the two sub-window crops (points in the reference box's frame +
    keep masks): resample both to `size`, append the timestamp (0 / 0.1) and prior-targetness (in the 1.25x box: 1 / 0 on
    the first tracked frame, 0.8 / 0.2 afterwards; 0.5 for the current frame) channels, stack previous over current, and
    give the BoxCloud of the previous half (zeros for the current half).  `draw(points, keep, size)` does the resampling."""
    dev = prev_local.device
    prev_pts = draw(prev_local, prev_keep, size)
    this_pts = draw(this_local, this_keep, size)
    canon = bx.Box(torch.zeros(3, device=dev), wlh, torch.eye(3, device=dev))
    half = torch.stack([wlh[1], wlh[0], wlh[2]]) * (1.25 / 2)
    inside = (prev_pts.abs() <= half).all(-1).float()          # nuscenes points_in_box: inclusive bounds
    mask_prev = inside if first_frame else inside * 0.6 + 0.2
    cols = lambda pts, t, m: torch.cat([pts, torch.full_like(pts[:, :1], t), m[:, None]], -1)
    stack = torch.cat([cols(prev_pts, 0.0, mask_prev), cols(this_pts, 0.1, torch.full_like(mask_prev, 0.5))], 0)
    data = {"points": stack[None]}
    if box_aware:
        bc = bx.point_to_box_distance(prev_pts, canon)
        data["candidate_bc"] = torch.cat([bc, torch.zeros_like(bc)], 0)[None]
    return data


class MotionBaseModel(BaseModel):
    """Base of the motion-centric models (models/base_model.py:250-303)."""

    def __init__(self, config, **kwargs):
        super().__init__(config, **kwargs)
        self.save_hyperparameters()

    def build_input_dict(self, sequence, frame_id, results_bbs, **kwargs):
        assert frame_id > 0, "no need to construct an input_dict at frame 0"
        cfg = self.config
        ref = _tbox(results_bbs[-1], self.device)
        crops = []
        for f in (sequence[frame_id - 1], sequence[frame_id]):
            local, keep = bx.subwindow(_points(f['pc'], self.device), ref, scale=cfg.bb_scale, offset=cfg.bb_offset)
            crops += [local, keep]

        def draw(points, keep, size):
            return regularize(points[keep], size, seed=1)[0]
        data = motion_input(*crops, ref.wlh, cfg.point_sample_size, frame_id == 1, getattr(cfg, 'box_aware', False), draw)
        return data, results_bbs[-1]
