"""Frames/s of the B=1 tracking frame on a synthetic tracklet:
  host    — the reference-shaped loop (models/base_model.py evaluate_one_sequence): per-frame crops with dynamic shapes,
            numpy index draw, box read-back every frame
  device  — tracking.DeviceTracker, eager (fixed shapes, no read-back)
  graph   — the same frame captured in one CUDA graph
usage: python tools/bench_track.py [--cfg BAT_Car.yaml] [--points 60000] [--frames 40]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from open3dsot_b200.config import load_config
from open3dsot_b200.datasets.synthetic import synthetic_sequence
from open3dsot_b200.models import get_model
from open3dsot_b200.tracking.device_tracker import DeviceTracker


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="BAT_Car.yaml")
    ap.add_argument("--points", type=int, default=60000)
    ap.add_argument("--frames", type=int, default=40)
    a = ap.parse_args()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = load_config(os.path.join(root, "cfgs", a.cfg), {})
    torch.manual_seed(0)
    net = get_model(cfg.net_model)(cfg).cuda().eval()
    seq = synthetic_sequence(n_frames=a.frames, n_points=a.points, seed=1)
    pts = [torch.tensor(f["pc"].points.T.copy(), device="cuda") for f in seq]

    net.evaluate_one_sequence(seq[:4])                       # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    net.evaluate_one_sequence(seq)
    torch.cuda.synchronize()
    host = (a.frames - 1) / (time.perf_counter() - t0)
    out = {"cfg": a.cfg, "points_per_scan": a.points, "frames": a.frames, "host_loop_fps": host}
    for name, graph in (("device_eager_fps", False), ("device_graph_fps", True)):
        trk = DeviceTracker(net, max_points=a.points, use_graph=graph)
        trk.reset(pts[0], seq[0]["3d_bbox"].to_tensor("cuda"))
        for i in range(1, 4):
            trk.step(pts[i])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(4, a.frames):
            trk.step(pts[i])
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        out[name] = (a.frames - 4) / max(wall, e0.elapsed_time(e1) * 1e-3)
        out[name.replace("fps", "ms_device")] = e0.elapsed_time(e1) / (a.frames - 4)
    print(out)


if __name__ == "__main__":
    main()
