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
        self.u_s = (torch.zeros(n, **f), torch.zeros(self.cfg.search_size, **f))
        self.u_t = (torch.zeros(2 * n, **f), torch.zeros(self.cfg.template_size, **f))
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
        local, keep, _ = bx.crop_and_center(self.scan, box, offset=self.cfg.model_bb_offset, scale=self.cfg.model_bb_scale)
        self.first_local.copy_(local)
        self.first_keep.copy_(keep & self.scan_valid)
        self.prev_scan.copy_(self.scan)
        self.prev_valid.copy_(self.scan_valid)
        self.frames = 1
        return self._box()

    # ------------------------------------------------------------------ one frame, fixed shapes
    def _inputs(self):
        cfg, box = self.cfg, self._box()
        # search area: current scan in the frame of the reference box (= previous result)
        b1 = bx.Box(box.center[None], box.wlh[None], box.rot[None])
        s_local, s_keep = bx.crop_in_box_frame(self.scan[None], b1, cfg.search_bb_scale, cfg.search_bb_offset)
        s_local, s_keep = s_local[0], s_keep[0]
        search, _ = resample(s_local, s_keep & self.scan_valid, cfg.search_size, u_perm=self.u_s[0], u_pick=self.u_s[1])
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
        data = {"template_points": template[None], "search_points": search[None]}
        if self.needs_bc:
            canon = bx.Box(torch.zeros_like(box.center), box.wlh, torch.eye(3, device=self.dev))
            data["points2cc_dist_t"] = bx.point_to_box_distance(template, canon)[None]
        return data

    def _frame(self):
        cfg = self.cfg
        with torch.no_grad():
            out = self.model(self._inputs())
            est = out["estimation_boxes"][0]                                   # (num_proposal, 5) or (4,)
            if est.dim() == 2:
                est = est.index_select(0, est[:, 4].argmax().reshape(1))[0, :4]    # (indexing by a 0-d tensor would sync)
            new = bx.offset_box(self._box(), est, degrees=cfg.degrees, use_z=cfg.use_z, limit_box=cfg.limit_box)
            self.box_c.copy_(new.center); self.box_r.copy_(new.rot)
            self.prev_scan.copy_(self.scan); self.prev_valid.copy_(self.scan_valid)

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
            snap = [t.clone() for t in (self.box_c, self.box_r, self.prev_scan, self.prev_valid)]
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._frame()                                                   # warm-up (allocations, autotuning)
            torch.cuda.current_stream().wait_stream(s)
            for t, v in zip((self.box_c, self.box_r, self.prev_scan, self.prev_valid), snap):
                t.copy_(v)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._frame()
            for t, v in zip((self.box_c, self.box_r, self.prev_scan, self.prev_valid), snap):
                t.copy_(v)
            self.graph.replay()
        else:
            self.graph.replay()
        self.frames += 1
        return self._box()
