"""The one-kernel fixed-shape resampling (csrc/resample.cu, `o3d_resample`) against the torch formulation it replaces on the device
(open3dsot_b200/tracking/sampling.py: radix top-k of random keys / cumsum + searchsorted), on the same uniform draws:
  * enough survivors: the SAME subset (the `size` smallest keys), emitted in ascending key order;
  * too few survivors: the same draws with replacement, element for element;
  * <= 2 survivors: the all-zero placeholder;
plus tied keys at the selection threshold, candidate counts that are not multiples of 4, a full-size scan and batches whose
clouds fall into different regimes."""
import pytest
import torch

from open3dsot_b200 import ops, runtime
from open3dsot_b200.tracking.sampling import resample_batched

pytestmark = pytest.mark.gpu


def _torch_path(points, keep, size, u_perm, u_pick):
    runtime.set_fused(False)
    try:
        return resample_batched(points, keep, size, u_perm, u_pick)
    finally:
        runtime.set_fused(True)


def _case(B, N, size, frac, seed, quant=None):
    g = torch.Generator().manual_seed(seed)
    pts = torch.randn(B, N, 3, generator=g).cuda()
    keep = (torch.rand(B, N, generator=g) < frac).cuda()
    u = torch.rand(B, N, generator=g)
    if quant:                                            # few distinct keys: ties everywhere, also at the threshold
        u = torch.floor(u * quant) / quant
    return pts, keep, u.cuda(), torch.rand(B, size, generator=g).cuda()


@pytest.mark.parametrize("B,N,size,frac,quant", [
    (1, 60000, 1024, 0.05, None),        # the tracker's search crop
    (1, 60000, 1024, 1.0, None),         # every candidate kept (worst case of the select)
    (2, 120000, 512, 0.02, None),        # template: first + previous crops concatenated
    (3, 4099, 512, 0.5, None),           # N % 4 != 0
    (2, 3000, 256, 0.4, 64),             # 64 distinct keys: heavy ties, the threshold falls inside a tie group
    (2, 2048, 2048, 1.0, None),          # n == size: everything, sorted by key
    (1, 777, 33, 0.3, None),             # size not a power of two
])
def test_without_replacement_selects_the_smallest_keys_in_key_order(B, N, size, frac, quant):
    pts, keep, u, up = _case(B, N, size, frac, seed=B * 1000 + size, quant=quant)
    out, src, n = ops.resample(pts, keep, size, u, up)
    out_t, src_t, n_t = _torch_path(pts, keep, size, u, up)
    assert torch.equal(n, n_t) and bool((n >= size).all())
    for b in range(B):
        assert bool(keep[b][src[b]].all()), "selected a dropped candidate"
        assert src[b].unique().numel() == size, "not distinct"
        ku, ku_t = u[b][src[b]], u[b][src_t[b]]
        assert torch.equal(ku.sort().values, ku_t.sort().values), "not the `size` smallest keys"
        if quant is None:
            assert torch.equal(src[b].sort().values, src_t[b].sort().values)
        else:                                            # ties at the threshold: ours takes them in index order
            thr = ku.max()
            tied = torch.nonzero(keep[b] & (u[b] == thr))[:, 0]
            mine = src[b][ku == thr].sort().values
            assert torch.equal(mine, tied[: mine.numel()])
        assert bool((ku[1:] >= ku[:-1]).all()), "not in ascending key order"
        assert torch.equal(out[b], pts[b][src[b]])


@pytest.mark.parametrize("B,N,size,frac", [(2, 5000, 1024, 0.05), (1, 60000, 512, 0.001), (3, 1001, 64, 0.02)])
def test_with_replacement_matches_the_torch_path_exactly(B, N, size, frac):
    pts, keep, u, up = _case(B, N, size, frac, seed=N + size)
    out, src, n = ops.resample(pts, keep, size, u, up)
    out_t, src_t, n_t = _torch_path(pts, keep, size, u, up)
    assert bool(((n > 2) & (n < size)).all()), "case is not in the with-replacement regime"
    assert torch.equal(n, n_t) and torch.equal(src, src_t) and torch.equal(out, out_t)


def test_placeholder_and_mixed_regimes_in_one_batch():
    B, N, size = 5, 3000, 128
    pts, keep, u, up = _case(B, N, size, 0.5, seed=9)
    keep[0] = False                                      # no survivor
    keep[1] = False; keep[1, 17] = True; keep[1, 2999] = True      # two survivors
    keep[2] = False; keep[2, 5:55] = True                # 50 < size: with replacement
    keep[3] = True                                       # everything
    out, src, n = ops.resample(pts, keep, size, u, up)
    out_t, src_t, n_t = _torch_path(pts, keep, size, u, up)
    assert n.tolist() == [0, 2, 50, N, int(keep[4].sum())]
    assert torch.equal(n, n_t)
    assert bool((out[:2] == 0).all()) and torch.equal(out[:3], out_t[:3]) and torch.equal(src[:3], src_t[:3])
    for b in (3, 4):
        assert torch.equal(src[b].sort().values, src_t[b].sort().values)
        assert torch.equal(out[b], pts[b][src[b]])


def test_resample_batched_routes_to_the_kernel_and_is_seed_reproducible():
    pts, keep, _, _ = _case(2, 8000, 512, 0.3, seed=4)
    g1 = torch.Generator(device="cuda").manual_seed(11)
    g2 = torch.Generator(device="cuda").manual_seed(11)
    a = resample_batched(pts, keep, 512, generator=g1)
    b = resample_batched(pts, keep, 512, generator=g2)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert bool(keep[0][a[1][0]].all()) and a[1][0].unique().numel() == 512
