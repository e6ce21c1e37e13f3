"""Cross-correlation front-end kernels (csrc/xcorr.cu) through the C ABI: BoxAware top-k (o3d_xcorr_boxaware_fwd) against
an exact statement of `cdist -> stable argsort -> [:k]` (models/head/xcorr.py:81-88) and the P2B cosine map
(o3d_xcorr_p2b_fwd / _bwd) against `F.cosine_similarity` and its autograd gradient (xcorr.py:37-38)."""
import pytest
import torch
import torch.nn.functional as F

from open3dsot_b200 import fused, ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,M,N,D,k", [(4, 64, 128, 9, 4), (2, 32, 64, 9, 4), (3, 12, 20, 9, 4), (1, 8, 5, 3, 8), (2, 64, 128, 9, 1)])
def test_boxaware_topk_exact(B, M, N, D, k):
    g = torch.Generator().manual_seed(B * 1000 + M)
    # coordinates on a 1/64 grid: every squared distance is exact in fp32, so ties are exact and the expected order is
    # unambiguous (ascending distance, equal distances in ascending template index)
    t = torch.randint(0, 48, (B, M, D), generator=g).float() / 64
    s = torch.randint(0, 48, (B, N, D), generator=g).float() / 64
    t[:, M // 2:M // 2 + 3] = t[:, 0:3]                                    # duplicated template points: exact ties
    d2 = ((t[:, :, None, :].double() - s[:, None, :, :].double()) ** 2).sum(-1)        # (B, M, N)
    want = torch.argsort(d2, dim=1, stable=True)[:, :k, :].transpose(1, 2).int()
    got = ops.boxaware_topk(t.cuda(), s.cuda(), k).cpu()
    assert torch.equal(got, want)


def test_boxaware_topk_random_matches_cdist_choice_up_to_near_ties():
    g = torch.Generator().manual_seed(3)
    t, s = torch.rand(8, 64, 9, generator=g), torch.rand(8, 128, 9, generator=g)
    got = ops.boxaware_topk(t.cuda(), s.cuda(), 4).cpu().long()
    dist = torch.cdist(t, s)                                                # the reference's formulation (matmul-based)
    want = torch.argsort(dist, dim=1, stable=True)[:, :4, :].transpose(1, 2)
    diff = got != want
    assert float(diff.float().mean()) < 2e-3
    if diff.any():                                                          # every disagreement is a near-tie in cdist's own terms
        b, n, kk = diff.nonzero(as_tuple=True)
        da, db = dist[b, got[b, n, kk], n], dist[b, want[b, n, kk], n]
        assert float((da - db).abs().max()) < 1e-5


@pytest.mark.parametrize("B,n1,n2,C", [(4, 64, 128, 256), (2, 16, 20, 16), (1, 100, 33, 40)])
def test_p2b_cosine_forward_backward(B, n1, n2, C):
    g = torch.Generator().manual_seed(C)
    t = torch.randn(B, n1, C, generator=g)
    s = torch.randn(B, n2, C, generator=g)
    t[0, 1] = 0.0                                                           # zero-norm row: the eps clamp
    s[0, 2] *= 1e-9
    td, sd = t.double().requires_grad_(True), s.double().requires_grad_(True)
    ref = F.cosine_similarity(td.transpose(1, 2).unsqueeze(-1), sd.transpose(1, 2).unsqueeze(2), dim=1)   # (B, n1, n2)
    w = torch.randn(B, n2, n1, generator=g)
    (ref.transpose(1, 2) * w.double()).sum().backward()
    tc, sc = t.cuda().requires_grad_(True), s.cuda().requires_grad_(True)
    sim = fused._P2BCosine.apply(tc, sc)                                    # (B, n2, n1)
    (sim * w.cuda()).sum().backward()
    assert float((sim.detach().cpu().double() - ref.detach().transpose(1, 2)).abs().max()) < 2e-6
    for a, b in ((tc.grad, td.grad), (sc.grad, sd.grad)):
        live = b.norm(dim=-1) < 1e6                                         # rows scaled by 1/eps excluded from the relative bound
        err = (a.cpu().double() - b)[live].norm() / b[live].norm()
        assert float(err) < 1e-5
