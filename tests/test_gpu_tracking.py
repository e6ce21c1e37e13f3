"""B=1 tracking inference on the GPU: the reference-shaped host loop and the graph-captured device tracker drive the
real BAT / P2B networks (eval mode, fused kernels) over a synthetic tracklet."""
import numpy as np
import pytest
import torch

from open3dsot_b200.config import load_config
from open3dsot_b200.datasets.synthetic import synthetic_sequence
from open3dsot_b200.models import get_model
from open3dsot_b200.tracking.device_tracker import DeviceTracker

pytestmark = pytest.mark.gpu


def _model(cfg_name):
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = load_config(os.path.join(root, "cfgs", cfg_name), {"up_axis": [0, 0, 1]})   # the synthetic tracklet is z-up
    torch.manual_seed(0)
    return cfg, get_model(cfg.net_model)(cfg).cuda().eval()


@pytest.mark.parametrize("cfg_name", ["BAT_Car.yaml", "P2B_Car.yaml"])
def test_host_loop_and_device_tracker_run_the_network(cfg_name):
    cfg, net = _model(cfg_name)
    seq = synthetic_sequence(n_frames=5, n_points=8000, seed=11)
    ious, dists, boxes = net.evaluate_one_sequence(seq)
    assert len(boxes) == 5 and ious[0] == pytest.approx(1.0) and all(np.isfinite(ious)) and all(np.isfinite(dists))
    for b in boxes[1:]:                                      # rigid update: orientation stays a rotation, size unchanged
        assert np.abs(b.rotation_matrix @ b.rotation_matrix.T - np.eye(3)).max() < 1e-5
        assert np.allclose(b.wlh, seq[0]["3d_bbox"].wlh)

    pts = [torch.tensor(f["pc"].points.T.copy(), device="cuda") for f in seq]
    res = {}
    for graph in (False, True):
        trk = DeviceTracker(net, max_points=8000, use_graph=graph)
        trk.reset(pts[0], seq[0]["3d_bbox"].to_tensor("cuda"))
        out = []
        for i in range(1, 5):
            b = trk.step(pts[i])
            out.append((b.center.clone(), b.rot.clone()))
        assert (trk.graph is not None) == graph
        res[graph] = out
    for (c, r) in res[False] + res[True]:
        assert torch.isfinite(c).all() and float((r @ r.t() - torch.eye(3, device="cuda")).abs().max()) < 1e-5
    # an untrained network's proposals are arbitrary but bounded by its own vote/offset scale: the box cannot fly away
    start = torch.tensor(seq[0]["3d_bbox"].center, device="cuda", dtype=torch.float32)
    for graph in (False, True):
        assert float((res[graph][-1][0] - start).norm()) < 50.0


@pytest.mark.parametrize("cfg_name", ["BAT_Car.yaml", "M2_track_kitti.yaml"])
def test_graph_replay_equals_eager_frames(cfg_name):
    """Same seed -> same resampling draws -> the captured frame and the eager frame produce the same boxes."""
    cfg, net = _model(cfg_name)
    seq = synthetic_sequence(n_frames=5, n_points=8000, seed=3)
    pts = [torch.tensor(f["pc"].points.T.copy(), device="cuda") for f in seq]
    tracks = []
    for graph in (False, True):
        trk = DeviceTracker(net, max_points=8000, use_graph=graph, seed=7)
        trk.reset(pts[0], seq[0]["3d_bbox"].to_tensor("cuda"))
        tracks.append([trk.step(p).center.clone() for p in pts[1:]])
    for a, b in zip(*tracks):
        assert float((a - b).abs().max()) < 1e-3


def test_crop_kernel_matches_tensor_formulation():
    """csrc/geometry.cu against boxes.crop_in_box_frame's tensor path (the one the CPU tests pin to the numpy restatement)."""
    from open3dsot_b200.tracking import boxes as bx
    g = torch.Generator().manual_seed(0)
    F, N, B = 5, 4099, 7
    scans = torch.randn(F, N, 3, generator=g) * 4
    count = torch.tensor([N, 4000, 17, 0, 2500])
    frame = torch.tensor([0, 1, 2, 3, 4, 1, 0])
    q, _ = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))
    q = q * torch.sign(torch.linalg.det(q))[:, None, None]
    box = bx.Box(torch.randn(B, 3, generator=g), torch.rand(B, 3, generator=g) * 3 + 1, q)
    want_l, want_k = bx.crop_in_box_frame(scans, box, 1.25, 2.0, frame, count)                   # CPU tensors: tensor path
    cu = bx.Box(*(t.cuda() for t in box))
    got_l, got_k = bx.crop_in_box_frame(scans.cuda(), cu, 1.25, 2.0, frame.cuda(), count.cuda())  # CUDA: fused kernel
    assert float((got_l.cpu() - want_l).abs().max()) < 1e-5
    half = torch.stack([box.wlh[:, 1], box.wlh[:, 0], box.wlh[:, 2]], -1) * 0.625 + 2.0
    edge = ((want_l.abs() - half[:, None, :]).abs() < 1e-4).any(-1)                               # rounding may flip these
    assert bool(((got_k.cpu() == want_k) | edge).all())
    assert int(got_k.cpu()[2].sum()) <= 17 and int(got_k.cpu()[3].sum()) == 0                    # padding rows never kept
    l1, k1 = bx.crop_in_box_frame(scans[:B].cuda() if F >= B else scans.cuda().repeat(2, 1, 1)[:B], cu, 1.0, 0.0)   # frame=None, count=None
    ref_l, ref_k = bx.crop_in_box_frame((scans.repeat(2, 1, 1))[:B], box, 1.0, 0.0)
    assert float((l1.cpu() - ref_l).abs().max()) < 1e-5


def test_m2track_frame_loop_runs_the_network():
    """Motion-centric model: MotionBaseModel.build_input_dict + M2-Track forward + box update over a synthetic tracklet."""
    cfg, net = _model("M2_track_kitti.yaml")
    seq = synthetic_sequence(n_frames=4, n_points=8000, seed=13)
    ious, dists, boxes = net.evaluate_one_sequence(seq)
    assert len(boxes) == 4 and ious[0] == pytest.approx(1.0) and all(np.isfinite(ious)) and all(np.isfinite(dists))
    for b in boxes[1:]:
        assert np.abs(b.rotation_matrix @ b.rotation_matrix.T - np.eye(3)).max() < 1e-5


@pytest.mark.parametrize("cfg_name,kind", [("BAT_Car.yaml", "siamese"), ("M2_track_kitti.yaml", "motion")])
def test_device_built_batches_train_the_network(cfg_name, kind):
    """Batches constructed on the device (graph-captured sampler, fused crop kernel) drive a training step."""
    from open3dsot_b200.datasets.device_sampler import DeviceMotionSampler, DeviceSiameseSampler
    cfg, net = _model(cfg_name)
    cfg.batch_size = 6
    net.train()
    tracklets = [synthetic_sequence(n_frames=5, n_points=6000, seed=40 + i, n_object=500) for i in range(3)]
    smp = (DeviceSiameseSampler if kind == "siamese" else DeviceMotionSampler)(tracklets, cfg, "cuda", seed=3)
    b1, v1 = smp.next_batch()
    first = {k: v.clone() for k, v in b1.items()}
    b2, v2 = smp.next_batch()                                    # replay of the captured construction: a new draw
    assert bool(v1.all()) and bool(v2.all())
    key = "search_points" if kind == "siamese" else "points"
    assert not torch.equal(first[key], b2[key])
    loss = net.training_step({k: v.clone() for k, v in b2.items()}, 0)
    loss.backward()
    assert torch.isfinite(loss) and all(torch.isfinite(p.grad).all() for p in net.parameters() if p.grad is not None)
