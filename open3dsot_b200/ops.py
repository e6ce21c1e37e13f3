"""Tensor-level entry points: torch tensors in, C-ABI calls on the current CUDA stream, torch tensors out.

torch is used for device memory and the stream only.  Argument checks follow upstream `pointnet2_ops`
(CHECK_CONTIGUOUS / CHECK_IS_FLOAT / CHECK_IS_INT / CHECK_CUDA -> RuntimeError; "CPU not supported").
"""
import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk_f(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (CPU not supported)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be a float tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")


def _chk_i(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (CPU not supported)")
    if t.dtype != torch.int32:
        raise RuntimeError(f"{name} must be an int tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")


LAUNCHES = 0  # kernels enqueued through the C ABI (bench.py reports it as gpu_launches)


def _call(name, *args):
    global LAUNCHES
    LAUNCHES += 1
    _lib.check(getattr(_lib.lib(), name)(*args), name)


# ------------------------------------------------------------------ the nine `_ext` entry points
def furthest_point_sampling(xyz, npoint):
    _chk_f(xyz, "xyz")
    B, N, _ = xyz.shape
    out = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
    _call("o3d_fps", xyz.data_ptr(), B, N, int(npoint), out.data_ptr(), _stream())
    return out


def gather_points(features, idx):
    _chk_f(features, "features"); _chk_i(idx, "idx")
    B, C, N = features.shape
    M = idx.shape[1]
    out = torch.empty(B, C, M, dtype=torch.float32, device=features.device)
    _call("o3d_gather", features.data_ptr(), idx.data_ptr(), B, C, N, M, out.data_ptr(), _stream())
    return out


def gather_points_grad(grad_out, idx, N):
    _chk_f(grad_out, "grad_out"); _chk_i(idx, "idx")
    B, C, M = grad_out.shape
    out = torch.zeros(B, C, N, dtype=torch.float32, device=grad_out.device)
    _call("o3d_gather_grad", grad_out.data_ptr(), idx.data_ptr(), B, C, int(N), M, out.data_ptr(), _stream())
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    _chk_f(new_xyz, "new_xyz"); _chk_f(xyz, "xyz")
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    out = torch.empty(B, M, nsample, dtype=torch.int32, device=xyz.device)
    _call("o3d_ball_query", new_xyz.data_ptr(), xyz.data_ptr(), B, N, M, float(radius), int(nsample), out.data_ptr(),
          _stream())
    return out


def group_points(features, idx):
    _chk_f(features, "features"); _chk_i(idx, "idx")
    B, C, N = features.shape
    _, M, S = idx.shape
    out = torch.empty(B, C, M, S, dtype=torch.float32, device=features.device)
    _call("o3d_group", features.data_ptr(), idx.data_ptr(), B, C, N, M, S, out.data_ptr(), _stream())
    return out


def group_points_grad(grad_out, idx, N):
    _chk_f(grad_out, "grad_out"); _chk_i(idx, "idx")
    B, C, M, S = grad_out.shape
    out = torch.zeros(B, C, N, dtype=torch.float32, device=grad_out.device)
    _call("o3d_group_grad", grad_out.data_ptr(), idx.data_ptr(), B, C, int(N), M, S, out.data_ptr(), _stream())
    return out


def three_nn(unknown, known):
    _chk_f(unknown, "unknown"); _chk_f(known, "known")
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = torch.empty(B, n, 3, dtype=torch.float32, device=unknown.device)
    idx = torch.empty(B, n, 3, dtype=torch.int32, device=unknown.device)
    _call("o3d_three_nn", unknown.data_ptr(), known.data_ptr(), B, n, m, dist2.data_ptr(), idx.data_ptr(), _stream())
    return dist2, idx


def three_interpolate(features, idx, weight):
    _chk_f(features, "features"); _chk_i(idx, "idx"); _chk_f(weight, "weight")
    B, c, m = features.shape
    n = idx.shape[1]
    out = torch.empty(B, c, n, dtype=torch.float32, device=features.device)
    _call("o3d_three_interpolate", features.data_ptr(), idx.data_ptr(), weight.data_ptr(), B, c, m, n, out.data_ptr(),
          _stream())
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    _chk_f(grad_out, "grad_out"); _chk_i(idx, "idx"); _chk_f(weight, "weight")
    B, c, n = grad_out.shape
    out = torch.zeros(B, c, int(m), dtype=torch.float32, device=grad_out.device)
    _call("o3d_three_interpolate_grad", grad_out.data_ptr(), idx.data_ptr(), weight.data_ptr(), B, c, n, int(m),
          out.data_ptr(), _stream())
    return out


# ------------------------------------------------------------------ fused supersets (channels-last)
def ballquery_group(xyz, new_xyz, feat_cl, radius, nsample, normalize_xyz=False, return_idx=True):
    """xyz (B,N,3), new_xyz (B,M,3), feat_cl (B,N,C)|None -> grouped (B,M,S,C+4) [feat | dx dy dz 0], idx (B,M,S)."""
    _chk_f(xyz, "xyz"); _chk_f(new_xyz, "new_xyz")
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    C = 0
    if feat_cl is not None:
        _chk_f(feat_cl, "feat_cl")
        C = feat_cl.shape[2]
    grouped = torch.empty(B, M, nsample, C + 4, dtype=torch.float32, device=xyz.device)
    idx = torch.empty(B, M, nsample, dtype=torch.int32, device=xyz.device) if return_idx else None
    _call("o3d_ballquery_group", xyz.data_ptr(), new_xyz.data_ptr(), feat_cl.data_ptr() if C else None, B, N, M, C,
          float(radius), int(nsample), int(bool(normalize_xyz)), idx.data_ptr() if return_idx else None,
          grouped.data_ptr(), _stream())
    return grouped, idx


def ballquery_group_grad(grad_grouped, idx, N, radius, normalize_xyz, need_feat=True, need_xyz=False,
                         need_new_xyz=False):
    _chk_f(grad_grouped, "grad_grouped"); _chk_i(idx, "idx")
    B, M, S, row = grad_grouped.shape
    C = row - 4
    dev = grad_grouped.device
    gf = torch.zeros(B, N, C, dtype=torch.float32, device=dev) if (need_feat and C) else None
    gx = torch.zeros(B, N, 3, dtype=torch.float32, device=dev) if need_xyz else None
    gn = torch.zeros(B, M, 3, dtype=torch.float32, device=dev) if need_new_xyz else None
    _call("o3d_ballquery_group_grad", grad_grouped.data_ptr(), idx.data_ptr(), B, int(N), M, C, S, float(radius),
          int(bool(normalize_xyz)), gf.data_ptr() if gf is not None else None,
          gx.data_ptr() if gx is not None else None, gn.data_ptr() if gn is not None else None, _stream())
    return gf, gx, gn


def three_nn_interpolate(unknown, known, known_feat_cl):
    """unknown (B,n,3), known (B,m,3), known_feat_cl (B,m,c) -> out_cl (B,n,c), idx (B,n,3), weight (B,n,3)."""
    _chk_f(unknown, "unknown"); _chk_f(known, "known"); _chk_f(known_feat_cl, "known_feat_cl")
    B, n, _ = unknown.shape
    m, c = known_feat_cl.shape[1], known_feat_cl.shape[2]
    out = torch.empty(B, n, c, dtype=torch.float32, device=unknown.device)
    idx = torch.empty(B, n, 3, dtype=torch.int32, device=unknown.device)
    w = torch.empty(B, n, 3, dtype=torch.float32, device=unknown.device)
    _call("o3d_three_nn_interpolate", unknown.data_ptr(), known.data_ptr(), known_feat_cl.data_ptr(), B, n, m, c,
          out.data_ptr(), idx.data_ptr(), w.data_ptr(), _stream())
    return out, idx, w


def three_nn_interpolate_grad(grad_out_cl, idx, weight, m):
    _chk_f(grad_out_cl, "grad_out_cl"); _chk_i(idx, "idx"); _chk_f(weight, "weight")
    B, n, c = grad_out_cl.shape
    g = torch.zeros(B, int(m), c, dtype=torch.float32, device=grad_out_cl.device)
    _call("o3d_three_nn_interpolate_grad", grad_out_cl.data_ptr(), idx.data_ptr(), weight.data_ptr(), B, n, int(m), c,
          g.data_ptr(), _stream())
    return g


# ------------------------------------------------------------------ cross-correlation front ends (models/head/xcorr.py)
def boxaware_topk(template_bc, search_bc, k):
    """template_bc (B,M,D), search_bc (B,N,D) -> idx (B,N,k) int32: nearest template box clouds per search point."""
    _chk_f(template_bc, "template_bc"); _chk_f(search_bc, "search_bc")
    B, M, D = template_bc.shape
    N = search_bc.shape[1]
    idx = torch.empty(B, N, int(k), dtype=torch.int32, device=search_bc.device)
    _call("o3d_xcorr_boxaware_fwd", template_bc.data_ptr(), search_bc.data_ptr(), B, M, N, D, int(k), idx.data_ptr(), _stream())
    return idx


def p2b_cosine(tfeat_cl, sfeat_cl, eps=1e-8):
    """tfeat_cl (B,n1,C), sfeat_cl (B,n2,C) -> sim (B,n2,n1), tnorm (B,n1), snorm (B,n2)."""
    _chk_f(tfeat_cl, "tfeat_cl"); _chk_f(sfeat_cl, "sfeat_cl")
    B, n1, C = tfeat_cl.shape
    n2 = sfeat_cl.shape[1]
    dev = tfeat_cl.device
    sim = torch.empty(B, n2, n1, dtype=torch.float32, device=dev)
    tn = torch.empty(B, n1, dtype=torch.float32, device=dev)
    sn = torch.empty(B, n2, dtype=torch.float32, device=dev)
    _call("o3d_xcorr_p2b_fwd", tfeat_cl.data_ptr(), sfeat_cl.data_ptr(), B, n1, n2, C, float(eps), sim.data_ptr(), tn.data_ptr(),
          sn.data_ptr(), _stream())
    return sim, tn, sn


def p2b_cosine_grad(dsim, sim, tfeat_cl, sfeat_cl, tn, sn, eps=1e-8, need_t=True, need_s=True):
    _chk_f(dsim, "dsim")
    B, n1, C = tfeat_cl.shape
    n2 = sfeat_cl.shape[1]
    dt = torch.empty_like(tfeat_cl) if need_t else None
    dsf = torch.empty_like(sfeat_cl) if need_s else None
    _call("o3d_xcorr_p2b_bwd", dsim.data_ptr(), sim.data_ptr(), tfeat_cl.data_ptr(), sfeat_cl.data_ptr(), tn.data_ptr(),
          sn.data_ptr(), B, n1, n2, C, float(eps), dt.data_ptr() if need_t else None, dsf.data_ptr() if need_s else None, _stream())
    return dt, dsf


# ------------------------------------------------------------------ box-frame crop (tracking loop / training sampler)
RESAMPLE_MAX_SIZE = 2048


def resample(points, keep, size, u_perm, u_pick):
    """Fixed-shape resampling in one kernel (csrc/resample.cu; semantics of tracking/sampling.py): points (B, N, 3) fp32 CUDA,
    keep (B, N) bool, u_perm (B, N) / u_pick (B, size) uniform [0, 1) -> out (B, size, 3), src (B, size) int64, n (B,) int64."""
    _chk_f(points, "points")
    B, N, _ = points.shape
    keep = keep.contiguous()
    u_perm, u_pick = u_perm.contiguous(), u_pick.contiguous()
    assert keep.dtype == torch.bool and keep.shape == (B, N) and u_perm.shape == (B, N) and u_pick.shape == (B, size)
    assert u_perm.dtype == torch.float32 and u_pick.dtype == torch.float32
    dev = points.device
    scratch = torch.empty(B, N, dtype=torch.int32, device=dev)
    out = torch.empty(B, size, 3, device=dev)
    src = torch.empty(B, size, dtype=torch.int64, device=dev)
    n = torch.empty(B, dtype=torch.int64, device=dev)
    _call("o3d_resample", points.data_ptr(), keep.data_ptr(), u_perm.data_ptr(), u_pick.data_ptr(), B, N, int(size),
          scratch.data_ptr(), out.data_ptr(), src.data_ptr(), n.data_ptr(), _stream())
    return out, src, n


def crop_box_frame(scans, center, rot, half, frame=None, count=None):
    """scans (F, N, 3) fp32 CUDA; center (B, 3), rot (B, 3, 3), half (B, 3); frame (B,) int64 picks a scan per sample
    (None: sample b reads scan b), count (F,) int64 = valid points per scan.  Returns local (B, N, 3), keep (B, N) bool."""
    _chk_f(scans, "scans")
    F, N, _ = scans.shape
    B = center.shape[0]
    center, rot, half = (t.contiguous().float() for t in (center, rot, half))
    local = torch.empty(B, N, 3, device=scans.device)
    keep = torch.empty(B, N, dtype=torch.bool, device=scans.device)
    _call("o3d_crop_box_frame", scans.data_ptr(), None if count is None else count.data_ptr(),
          None if frame is None else frame.data_ptr(), center.data_ptr(), rot.data_ptr(), half.data_ptr(), B, N,
          local.data_ptr(), keep.data_ptr(), _stream())
    return local, keep
