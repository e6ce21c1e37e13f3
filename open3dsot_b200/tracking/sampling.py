"""Fixed-shape resampling of a masked point set — the device-side form of points_utils.regularize_pc (:24-40).

The reference draws `numpy.random.default_rng(seed=1).choice(n, size, replace=size > n)` on the host, which needs the
number of surviving points on the host (a device->host sync per crop) and a host RNG.  Here the candidates stay a
fixed-size array plus a keep-mask and the draw is expressed with sorts and gathers of fixed shape, so a whole frame
(crop, resample, model, box update) can be captured in one CUDA graph:
  * n >= size : `size` distinct survivors, uniformly at random (the `size` smallest of per-point random keys);
  * 2 < n < size : `size` draws with replacement;
  * n <= 2 : all-zero cloud (the reference's "too few points" placeholder).
The subset differs from the numpy Generator's (same distribution); tests that compare against the host restatement pass
the oracle's indices through `indices=`."""
import torch

from .. import ops, runtime


def resample(points, keep, size, generator=None, indices=None, u_perm=None, u_pick=None):
    """points (N, 3), keep (N,) bool -> (size, 3) points, (size,) source indices (into `points`).
    `u_perm` (N,) / `u_pick` (size,): uniform [0, 1) draws supplied by the caller (static buffers refreshed outside a
    captured graph, which keeps a replayed frame reproducible from a seed); drawn here when absent."""
    if indices is not None:                                   # explicit indices INTO THE SURVIVORS, in their order
        order = torch.nonzero(keep, as_tuple=False)[:, 0]
        src = order[indices]
        return points[src], src
    out, src, _ = resample_batched(points[None], keep[None], size, None if u_perm is None else u_perm[None],
                                   None if u_pick is None else u_pick[None], generator)
    return out[0], src[0]


def resample_batched(points, keep, size, u_perm=None, u_pick=None, generator=None):
    """Batched form: points (B, N, 3), keep (B, N) -> (B, size, 3) points, (B, size) source indices, survivor counts (B,).
    Without replacement = the `size` smallest random keys among the survivors (a radix select, not a full sort);
    with replacement = random ranks mapped to survivors through the running count of the keep-mask."""
    B, n_all = keep.shape
    dev = points.device
    u = torch.rand(B, n_all, device=dev, generator=generator) if u_perm is None else u_perm
    if points.is_cuda and points.dtype == torch.float32 and points.shape[-1] == 3 and 3 <= size <= ops.RESAMPLE_MAX_SIZE \
            and runtime.fused_enabled():
        # one kernel (csrc/resample.cu): ordered compaction, radix select of the size-th smallest key, sort, gather
        up = torch.rand(B, size, device=dev, generator=generator) if u_pick is None else u_pick
        return ops.resample(points.contiguous(), keep, size, u, up)
    n = keep.sum(1)
    key = torch.where(keep, u, torch.full_like(u, 2.0))
    k = min(size, n_all)
    wo = torch.topk(key, k, dim=1, largest=False, sorted=False).indices
    if k < size:
        wo = torch.cat([wo, wo.new_zeros(B, size - k)], 1)
    up = torch.rand(B, size, device=dev, generator=generator) if u_pick is None else u_pick
    rank = torch.minimum((up * n[:, None]).long(), torch.clamp(n - 1, min=0)[:, None])
    running = torch.cumsum(keep, 1, dtype=torch.int32)
    w = torch.searchsorted(running, (rank + 1).to(torch.int32)).clamp_(max=n_all - 1)
    src = torch.where((n >= size)[:, None], wo, w)
    out = torch.gather(points, 1, src[..., None].expand(-1, -1, points.shape[-1]))
    return torch.where((n > 2)[:, None, None], out, torch.zeros_like(out)), src, n
