"""B=1 tracking inference with the whole frame on the device.

The reference's frame loop (models/base_model.py:44-117, :166-247) crops and resamples on the host with numpy /
pyquaternion, uploads two small clouds, runs the network, downloads the proposals (`.cpu().numpy()`, :48) and updates the
box on the host.  Here a frame is: scan already on the device -> search-area crop in the previous box's frame
(generate_search_area :197-218) -> template = first-frame crop + previous-frame crop (generate_template :166-195,
shape_aggregation 'firstandprevious' / 'first' / 'previous') -> fixed-shape resampling (sampling.py) -> BoxCloud of the
template (bat.py:41-55) -> network in eval mode on the fused kernels -> best proposal -> box update (getOffsetBB).
Every tensor has a static shape, nothing is read back, so the frame is captured once in a CUDA graph and replayed."""
import torch

from . import boxes as bx
from .sampling import resample


class DeviceTracker:
    def __init__(self, model, max_points, use_graph=True, seed=1):
        self.model = model.eval()
        self.cfg = model.config
        self.dev = next(model.parameters()).device
        self.max_points = int(max_points)
        self.use_graph = bool(use_graph) and self.dev.type == "cuda"
        self.gen = torch.Generator(device=self.dev).manual_seed(seed)
        self.needs_bc = hasattr(model, "mlp_bc")            # BAT consumes the template BoxCloud
        self.motion = "point_sample_size" in self.cfg and not hasattr(model, "backbone")   # M2-Track style two-frame input
        f = dict(device=self.dev, dtype=torch.float32)
        n = self.max_points
        # static buffers (graph inputs / state)
        self.scan = torch.zeros(n, 3, **f)
        self.scan_valid = torch.zeros(n, dtype=torch.bool, device=self.dev)
        self.first_local = torch.zeros(n, 3, **f)           # first-frame object crop, canonical frame
        self.first_keep = torch.zeros(n, dtype=torch.bool, device=self.dev)
        self.prev_scan = torch.zeros(n, 3, **f)
        self.prev_valid = torch.zeros(n, dtype=torch.bool, device=self.dev)
        self.box_c = torch.zeros(3, **f)
        self.box_s = torch.ones(3, **f)
        self.box_r = torch.eye(3, **f)
        # uniform draws of the two resamplings: refreshed per frame OUTSIDE the captured graph (seeded, reproducible)
        size_s = self.cfg.point_sample_size if self.motion else self.cfg.search_size
        size_t = self.cfg.point_sample_size if self.motion else self.cfg.template_size
        self.u_s = (torch.zeros(n, **f), torch.zeros(size_s, **f))
        self.u_t = (torch.zeros(2 * n, **f), torch.zeros(size_t, **f))
        self.first_flag = torch.ones((), **f)               # 1 on the first tracked frame (prior box = ground truth), then 0
        self.graph = None
        self.frames = 0

    # ------------------------------------------------------------------ state helpers
    def _box(self):
        return bx.Box(self.box_c, self.box_s, self.box_r)

    def _load_scan(self, points, n_valid=None):
        n = points.shape[0]
        if n > self.max_points:
            raise ValueError(f"scan has {n} points, tracker was built for {self.max_points}")
        self.scan.zero_()
        self.scan[:n].copy_(points)
        self.scan_valid.zero_()
        self.scan_valid[: (n if n_valid is None else int(n_valid))] = True

    def reset(self, points, box: bx.Box):
        """First frame: remember the object crop (cropAndCenterPC of the first box) and the box itself."""
        self._load_scan(points)
        box = box.to(self.dev)
        self.box_c.copy_(box.center); self.box_s.copy_(box.wlh); self.box_r.copy_(box.rot)
        if not self.motion:
            local, keep, _ = bx.crop_and_center(self.scan, box, offset=self.cfg.model_bb_offset, scale=self.cfg.model_bb_scale)
            self.first_local.copy_(local)
            self.first_keep.copy_(keep & self.scan_valid)
        self.first_flag.fill_(1.0)
        self.prev_scan.copy_(self.scan)
        self.prev_valid.copy_(self.scan_valid)
        self.frames = 1
        return self._box()

    # ------------------------------------------------------------------ one frame, fixed shapes
    def _inputs_motion(self):
        """MotionBaseModel.build_input_dict (models/base_model.py:255-303) on static buffers."""
        cfg, box = self.cfg, self._box()
        b1 = bx.Box(box.center[None], box.wlh[None], box.rot[None])
        n = cfg.point_sample_size
        p_local, p_keep = bx.crop_in_box_frame(self.prev_scan[None], b1, cfg.bb_scale, cfg.bb_offset)
        t_local, t_keep = bx.crop_in_box_frame(self.scan[None], b1, cfg.bb_scale, cfg.bb_offset)
        prev_pts, _ = resample(p_local[0], p_keep[0] & self.prev_valid, n, u_perm=self.u_t[0][: self.max_points], u_pick=self.u_t[1])
        this_pts, _ = resample(t_local[0], t_keep[0] & self.scan_valid, n, u_perm=self.u_s[0], u_pick=self.u_s[1])
        half = torch.stack([box.wlh[1], box.wlh[0], box.wlh[2]]) * (1.25 / 2)
        inside = (prev_pts.abs() <= half).all(-1).float()
        f = self.first_flag
        mask_prev = inside * (0.6 + 0.4 * f) + 0.2 * (1 - f)               # 1 / 0 on the first frame, 0.8 / 0.2 afterwards
        col = lambda pts, t, m: torch.cat([pts, torch.full_like(pts[:, :1], t), m[:, None]], -1)
        data = {"points": torch.cat([col(prev_pts, 0.0, mask_prev), col(this_pts, 0.1, torch.full_like(mask_prev, 0.5))], 0)[None]}
        if getattr(cfg, "box_aware", False):
            canon = bx.Box(torch.zeros_like(box.center), box.wlh, torch.eye(3, device=self.dev))
            bc = bx.point_to_box_distance(prev_pts, canon)
            data["candidate_bc"] = torch.cat([bc, torch.zeros_like(bc)], 0)[None]
        return data

    def _inputs(self):
        if self.motion:
            return self._inputs_motion()
        cfg, box = self.cfg, self._box()
        b1 = bx.Box(box.center[None], box.wlh[None], box.rot[None])

        def build_template():
            # template: first-frame crop (+ previous-frame crop around the previous result)
            mode = cfg.shape_aggregation.upper()
            p_local, p_keep = bx.crop_in_box_frame(self.prev_scan[None], b1, cfg.model_bb_scale, cfg.model_bb_offset)
            p_local, p_keep = p_local[0], p_keep[0] & self.prev_valid
            if "FIRSTANDPREVIOUS" in mode:
                cand, keep = torch.cat([self.first_local, p_local]), torch.cat([self.first_keep, p_keep])
            elif "FIRST" in mode:
                cand, keep = self.first_local, self.first_keep
            elif "PREVIOUS" in mode:
                cand, keep = p_local, p_keep
            else:
                raise NotImplementedError(f"shape_aggregation '{cfg.shape_aggregation}' needs every past frame on the device")
            template, _ = resample(cand, keep, cfg.template_size, u_perm=self.u_t[0][: cand.shape[0]], u_pick=self.u_t[1])
            bc = None
            if self.needs_bc:
                canon = bx.Box(torch.zeros_like(box.center), box.wlh, torch.eye(3, device=self.dev))
                bc = bx.point_to_box_distance(template, canon)[None]
            return template[None], bc

        from .. import fused
        # the two crops + draws are independent: the template's run on the side stream (a parallel branch of the frame's graph)
        overlap = fused.branch_overlap(self.scan)
        join = fused.run_ahead(build_template) if overlap else None
        # search area: current scan in the frame of the reference box (= previous result)
        s_local, s_keep = bx.crop_in_box_frame(self.scan[None], b1, cfg.search_bb_scale, cfg.search_bb_offset)
        s_local, s_keep = s_local[0], s_keep[0]
        search, _ = resample(s_local, s_keep & self.scan_valid, cfg.search_size, u_perm=self.u_s[0], u_pick=self.u_s[1])
        template, bc = join() if overlap else build_template()
        data = {"template_points": template, "search_points": search[None]}
        if bc is not None:
            data["points2cc_dist_t"] = bc
        return data

    def _frame(self):
        cfg = self.cfg
        from .. import runtime
        # the tracker never changes the weights: eval-mode stacks pack them / fold their BatchNorm once, not per frame
        with torch.no_grad(), runtime.static_weights_scope():
            out = self.model(self._inputs())
            est = out["estimation_boxes"][0]                                   # (num_proposal, 5) or (4,)
            if est.dim() == 2:
                est = est.index_select(0, est[:, 4].argmax().reshape(1))[0, :4]    # (indexing by a 0-d tensor would sync)
            new = bx.offset_box(self._box(), est, degrees=cfg.degrees, use_z=cfg.use_z, limit_box=cfg.limit_box)
            self.box_c.copy_(new.center); self.box_r.copy_(new.rot)
            self.prev_scan.copy_(self.scan); self.prev_valid.copy_(self.scan_valid)
            self.first_flag.zero_()

    def step(self, points, n_valid=None):
        """Next frame: `points` (n, 3) device tensor.  Returns the tracked box (views of the tracker's state buffers)."""
        if self.frames == 0:
            raise RuntimeError("call reset() with the first frame and its box before step()")
        self._load_scan(points, n_valid)
        for u in self.u_s + self.u_t:
            u.uniform_(generator=self.gen)
        if not self.use_graph:
            self._frame()
        elif self.graph is None:
            state = (self.box_c, self.box_r, self.prev_scan, self.prev_valid, self.first_flag)
            snap = [t.clone() for t in state]
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._frame()                                                   # warm-up (allocations, autotuning)
            torch.cuda.current_stream().wait_stream(s)
            for t, v in zip(state, snap):
                t.copy_(v)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._frame()
            for t, v in zip(state, snap):
                t.copy_(v)
            self.graph.replay()
        else:
            self.graph.replay()
        self.frames += 1
        return self._box()
