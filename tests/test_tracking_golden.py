"""SURVEY.md §8f ranks 2-3 pinned to the reference: tests/golden/ref_tracking.npz was produced by the reference's own, unmodified
`datasets/points_utils.py` / `data_classes.py` / `sampler.py` (tests/golden/make_golden_tracking.py).  Held to it here:

  * the numpy oracle (oracle/tracking_ref.py) — function by function and the whole `siamese_processing`, on the recorded draws;
  * the product's tensor geometry (open3dsot_b200/tracking/boxes.py, datasets/device_sampler.py) on CPU tensors;
  * [-m gpu] the same product code on CUDA tensors, i.e. through csrc/geometry.cu (`o3d_crop_box_frame`) and the device sampler.
"""
import os

import numpy as np
import pytest
import torch

from oracle import tracking_ref as R
from open3dsot_b200.config import load_config
from open3dsot_b200.datasets.device_sampler import DeviceTracklets, siamese_batch
from open3dsot_b200.datasets.synthetic import synthetic_sequence
from open3dsot_b200.tracking import boxes as bx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "ref_tracking.npz"))


def _tracklets():
    return [synthetic_sequence(n_frames=int(G["meta.n_frames"]), n_points=int(G["meta.n_points"]), seed=int(s),
                               n_object=int(G["meta.n_object"])) for s in G["meta.tracklet_seeds"]]


def gbox(key):
    return R.Box(G[key + ".center"], G[key + ".wlh"], G[key + ".rot"])


def same_box(b, key, tol=1e-9):
    return (np.abs(b.center - G[key + ".center"]).max() < tol and np.abs(b.wlh - G[key + ".wlh"]).max() < tol
            and np.abs(b.rot - G[key + ".rot"]).max() < tol)


def rows_equal_as_sets(a, b, tol):
    """(n,3) point sets equal up to order"""
    if a.shape != b.shape:
        return False
    ka, kb = np.lexsort(np.round(a / (50 * tol)).T[::-1]), np.lexsort(np.round(b / (50 * tol)).T[::-1])
    if np.abs(a[ka] - b[kb]).max() < tol:
        return True
    d = np.abs(a[:, None, :] - b[None, :, :]).max(-1)      # fall back to a nearest-row match (rounding at a bucket edge)
    return d.min(1).max() < tol and d.min(0).max() < tol


# ---------------------------------------------------------------------------------------------- the numpy oracle
def test_oracle_functions_match_reference_outputs():
    frames = [f for t in _tracklets() for f in t]
    p1 = frames[1]["pc"].points.astype(np.float64)
    p0 = frames[0]["pc"].points.astype(np.float64)
    b0 = R.Box(frames[0]["3d_bbox"].center, frames[0]["3d_bbox"].wlh, frames[0]["3d_bbox"].rotation_matrix)
    box = gbox("box_in")
    for i in range(4):
        a = G[f"offset{i}.args"]
        n = len(a) - 3
        nb = R.get_offset_bb(box, a[:n], degrees=bool(a[n]), use_z=bool(a[n + 1]), limit_box=bool(a[n + 2]), rand=tuple(G[f"offset{i}.rand"]))
        assert same_box(nb, f"offset{i}.out"), i
    moved = gbox("box_moved")
    assert same_box(R.get_offset_bb(box, np.array([0.3, -0.2, 4.0]), degrees=True, limit_box=False), "box_moved")
    assert np.allclose(R.generate_subwindow(p1, moved, 1.25, 2), G["subwindow.points"], atol=1e-9)
    mp, mb = R.get_model([p0, p1], [b0, moved], offset=0, scale=1.25)
    assert np.allclose(mp, G["model.points"], atol=1e-9) and same_box(mb, "model.box")
    cp, cb = R.crop_and_center_pc(p1, moved, offset=0.5, scale=1.1)
    assert np.allclose(cp, G["cropcenter.points"], atol=1e-9) and same_box(cb, "cropcenter.box")
    reg, idx = R.regularize_pc(G["subwindow.points"].T, 1024, seed=1)
    assert np.array_equal(idx, G["regularize.idx"]) and np.array_equal(reg, G["regularize.points"])
    reg2, idx2 = R.regularize_pc(G["model.points"].T, 512, seed=1)
    assert np.array_equal(idx2, G["regularize_up.idx"])
    assert np.allclose(R.get_point_to_box_distance(reg2, mb), G["boxcloud"], atol=1e-9)
    tb = R.transform_box(R.Box(frames[1]["3d_bbox"].center, frames[1]["3d_bbox"].wlh, frames[1]["3d_bbox"].rotation_matrix), moved)
    assert same_box(tb, "transform_box")
    assert np.array_equal(R.get_in_box_mask(G["subwindow.points"], tb), G["in_box_mask"])


def _siamese_cases():
    return [(tag, s) for tag in ("bat", "p2b") for s in range(6)]


def _fr(f):
    b = f["3d_bbox"]
    return f["pc"].points.astype(np.float64), R.Box(b.center, b.wlh, b.rotation_matrix)


def _lim(v):
    return tuple(0.0 if np.isnan(x) else float(x) for x in v)


@pytest.mark.parametrize("tag,s", _siamese_cases())
def test_oracle_siamese_processing_matches_reference(tag, s):
    cfg = load_config(os.path.join(ROOT, "cfgs", "BAT_Car.yaml" if tag == "bat" else "P2B_Car.yaml"))
    frames = [f for t in _tracklets() for f in t]
    key = f"siamese.{tag}.{s}"
    i_first, i_prev, k, cand = (int(v) for v in G[key + ".frames"])
    out = R.siamese_processing(_fr(frames[i_first]), _fr(frames[i_prev]), _fr(frames[k]), cand, cfg, G[key + ".template_offset"],
                               G[key + ".search_offset"], idx_t=G[key + ".idx_t"], idx_s=G[key + ".idx_s"],
                               limit_rand_t=_lim(G[key + ".limit_rand_t"]), limit_rand_s=_lim(G[key + ".limit_rand_s"]))
    for name in ("template_points", "search_points", "box_label", "bbox_size", "seg_label") + (
            ("points2cc_dist_t", "points2cc_dist_s") if tag == "bat" else ()):
        want = G[f"{key}.out.{name}"]
        assert out[name].shape == want.shape and np.abs(out[name].astype(np.float64) - want).max() < 1e-5, name


# ---------------------------------------------------------------------------------------------- the product (tensor geometry)
def _tbox(key, dev):
    return bx.make_box(torch.tensor(G[key + ".center"], dtype=torch.float32, device=dev)[None],
                       torch.tensor(G[key + ".wlh"], dtype=torch.float32, device=dev)[None],
                       rot=torch.tensor(G[key + ".rot"], dtype=torch.float32, device=dev)[None])


def _product_geometry(dev):
    frames = [f for t in _tracklets() for f in t]
    p1 = torch.tensor(frames[1]["pc"].points.T.copy(), dtype=torch.float32, device=dev)[None]       # (1, N, 3)
    moved, box = _tbox("box_moved", dev), _tbox("box_in", dev)
    for i in range(4):
        a = G[f"offset{i}.args"]
        n = len(a) - 3
        nb = bx.offset_box(box, torch.tensor(a[:n], dtype=torch.float32, device=dev)[None], degrees=bool(a[n]), use_z=bool(a[n + 1]),
                           limit_box=bool(a[n + 2]), rand=torch.tensor(G[f"offset{i}.rand"], dtype=torch.float32, device=dev)[None])
        assert np.abs(nb.center[0].cpu().numpy() - G[f"offset{i}.out.center"]).max() < 1e-5, i
        assert np.abs(nb.rot[0].cpu().numpy() - G[f"offset{i}.out.rot"]).max() < 1e-5, i
    local, keep = bx.subwindow(p1, moved, 1.25, 2.0)
    got = local[0][keep[0]].cpu().numpy().astype(np.float64)
    assert rows_equal_as_sets(got, G["subwindow.points"].T, 2e-5)
    lc, kc, canon = bx.crop_and_center(p1, moved, offset=0.5, scale=1.1)
    assert rows_equal_as_sets(lc[0][kc[0]].cpu().numpy().astype(np.float64), G["cropcenter.points"].T, 2e-5)
    assert np.abs(canon.center[0].cpu().numpy() - G["cropcenter.box.center"]).max() < 1e-5
    pts = torch.tensor(G["regularize_up.points"], dtype=torch.float32, device=dev)[None]
    bc = bx.point_to_box_distance(pts, _tbox("model.box", dev))
    assert np.abs(bc[0].cpu().numpy() - G["boxcloud"]).max() < 2e-5
    # the fused crop kernel / its tensor formulation on a stack of scans with per-sample boxes
    scans = torch.stack([torch.tensor(f["pc"].points.T.copy(), dtype=torch.float32, device=dev) for f in frames[:4]])
    lk, kk = bx.crop_in_box_frame(scans, bx.make_box(moved.center.expand(2, -1), moved.wlh.expand(2, -1), rot=moved.rot.expand(2, -1, -1)),
                                  1.25, 2.0, frame=torch.tensor([1, 1], device=dev))
    assert rows_equal_as_sets(lk[0][kk[0]].cpu().numpy().astype(np.float64), G["subwindow.points"].T, 2e-5)


def _product_sampler(dev, tag):
    cfg = load_config(os.path.join(ROOT, "cfgs", "BAT_Car.yaml" if tag == "bat" else "P2B_Car.yaml"))
    data = DeviceTracklets(_tracklets(), dev)
    keys = [f"siamese.{tag}.{s}" for s in range(6)]
    fr = np.stack([G[k + ".frames"] for k in keys])
    deg = 5.0 if cfg.degrees else np.deg2rad(5.0)
    t = lambda name, f=lambda v: v: torch.tensor(np.stack([f(G[k + name]) for k in keys]), dtype=torch.float32, device=dev)   # noqa: E731
    nz = lambda v: np.nan_to_num(v, nan=0.0)                                                                                    # noqa: E731
    # the device sampler draws N(0, 1) and scales by sqrt(diag(1, 1, 5 deg)) itself: hand it the reference's sample un-scaled
    draws = {"template_offset": t(".template_offset"), "search_offset": t(".search_offset", lambda v: v / np.sqrt([1.0, 1.0, deg])),
             "limit_rand_t": t(".limit_rand_t", nz), "limit_rand_s": t(".limit_rand_s", nz)}
    batch, valid = siamese_batch(data, cfg, torch.tensor(fr[:, 2], device=dev), torch.tensor(fr[:, 3], device=dev), draws=draws)
    assert bool(valid.all())
    for b, k in enumerate(keys):
        want = {n: G[f"{k}.out.{n}"] for n in ("template_points", "search_points", "box_label", "bbox_size", "seg_label")}
        assert np.abs(batch["box_label"][b].cpu().numpy() - want["box_label"]).max() < 1e-4, k
        assert np.abs(batch["bbox_size"][b].cpu().numpy() - want["bbox_size"]).max() < 1e-6
        # survivors: the reference's crops before resampling = its resampled clouds' unique rows (it resamples by index)
        # survivor counts: the reference's index draws range over its own crops, so they bound the counts from below
        n_t, n_s = int(batch["_n_template"][b]), int(batch["_n_search"][b])
        assert n_t > int(G[k + ".idx_t"].max()) and n_s > int(G[k + ".idx_s"].max()), k
        tp, sp = batch["template_points"][b].cpu().numpy(), batch["search_points"][b].cpu().numpy()
        # every point the sampler emits is one of the reference's survivors, with the reference's label and (BAT) box cloud
        ref_s = want["search_points"].astype(np.float64)
        d = np.abs(sp[:, None, :].astype(np.float64) - ref_s[None, :, :]).max(-1)
        hit = d.min(1) < 3e-5
        # the reference drew 1024 of n_s survivors; a sampler point that the reference did not draw cannot be matched: only
        # matched points are label-checked, and at least half must match (both draw ~uniformly from the same survivors)
        assert hit.mean() > 0.5, (k, hit.mean())
        lab = want["seg_label"][d.argmin(1)]
        assert np.array_equal(batch["seg_label"][b].cpu().numpy()[hit] > 0.5, lab[hit] > 0.5), k
        ref_t = want["template_points"].astype(np.float64)
        dt = np.abs(tp[:, None, :].astype(np.float64) - ref_t[None, :, :]).max(-1)
        assert (dt.min(1) < 3e-5).mean() > 0.5, k
        if tag == "bat":
            bc = G[f"{k}.out.points2cc_dist_s"][d.argmin(1)]
            assert np.abs(batch["points2cc_dist_s"][b].cpu().numpy()[hit] - bc[hit]).max() < 1e-4, k


def test_product_geometry_matches_reference_on_cpu():
    _product_geometry("cpu")


@pytest.mark.parametrize("tag", ["bat", "p2b"])
def test_product_sampler_matches_reference_on_cpu(tag):
    _product_sampler("cpu", tag)


@pytest.mark.gpu
def test_product_geometry_matches_reference_on_cuda():
    _product_geometry("cuda")


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["bat", "p2b"])
def test_product_sampler_matches_reference_on_cuda(tag):
    _product_sampler("cuda", tag)
