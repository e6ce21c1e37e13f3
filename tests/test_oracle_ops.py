"""Known-answer tests for the C oracle of the nine `pointnet2_ops._ext` ops (SURVEY.md §2.3 / §8c "golden
vectors to create"): the semantics are pinned by hand-derived answers, not by another implementation."""
import numpy as np
import torch

from oracle import ops


def test_opt_n_threads_rule():
    assert [ops.opt_n_threads(n) for n in (1, 2, 3, 100, 511, 512, 1024, 5000)] == [1, 2, 2, 64, 256, 512, 512, 512]


def test_fps_starts_at_zero_and_picks_farthest():
    xyz = torch.tensor([[[1., 0, 0], [1.1, 0, 0], [5., 0, 0], [3., 0, 0]]])
    assert ops.furthest_point_sampling(xyz, 3).tolist() == [[0, 2, 3]]


def test_fps_skips_near_origin_points():
    # point 1 is the farthest from point 0 but has |p|^2 <= 1e-3 -> never selected
    xyz = torch.tensor([[[4., 0, 0], [0.01, 0.01, 0.01], [3., 0, 0], [1., 0, 0]]])
    assert ops.furthest_point_sampling(xyz, 3).tolist() == [[0, 3, 2]]


def test_fps_all_zero_cloud_returns_zeros():
    xyz = torch.zeros(2, 16, 3)  # regularize_pc zero-fill (datasets/points_utils.py:39)
    assert ops.furthest_point_sampling(xyz, 5).eq(0).all()


def test_fps_tie_order_is_bit_reversed_thread_id():
    # N=8 -> block 8, every thread owns one point.  Points 1..7 are all at distance 1 from point 0:
    # the tree keeps the left operand, so the winner is the smallest bit-reversed tid among {1..7}:
    # bitrev3: 1->4, 2->2, 3->6, 4->1, 5->5, 6->3, 7->7  => index 4 wins, not index 1.
    dirs = torch.tensor([[1., 0, 0], [-1., 0, 0], [0, 1., 0], [0, -1., 0], [0, 0, 1.], [0, 0, -1.], [1., 0, 0]])
    centre = torch.tensor([[10., 10., 10.]])
    xyz = torch.cat([centre, centre + dirs]).unsqueeze(0).contiguous()
    assert ops.furthest_point_sampling(xyz, 2).tolist() == [[0, 4]]


def test_fps_tie_within_thread_prefers_lower_index():
    # N=5 -> block 4: thread 0 owns k=0 and k=4, k=4 duplicates k=2's distance... build: points 1,2,3 and 4
    # equidistant from 0; tree order among tids {1,2,3} is bitrev2: 1->2, 2->1, 3->3 -> tid 2 (k=2) wins
    # over k=4 (tid 0) only if distances tie: tid 0's best is k=4 (k=0 has d=0). Left operand tid0 wins ties.
    c = torch.tensor([[5., 5., 5.]])
    xyz = torch.cat([c, c + torch.tensor([[1., 0, 0], [0, 1., 0], [0, 0, 1.], [-1., 0, 0]])]).unsqueeze(0).contiguous()
    assert ops.furthest_point_sampling(xyz, 2).tolist() == [[0, 4]]


def test_ball_query_order_padding_strictness_and_empty():
    xyz = torch.tensor([[[0., 0, 0], [0.5, 0, 0], [1.0, 0, 0], [0.2, 0, 0], [0.9, 0, 0], [100., 0, 0]]])
    new_xyz = torch.tensor([[[0., 0, 0], [50., 0, 0], [100., 0, 0]]])
    idx = ops.ball_query(new_xyz, xyz, 1.0, 4)
    # centre 0: d2<1 strictly -> k=0,1,3,4 (k=2 lies exactly on the radius -> excluded), ascending order
    assert idx[0, 0].tolist() == [0, 1, 3, 4]
    # centre 1: nothing in range -> zeros ; centre 2: one hit (k=5) -> replicated
    assert idx[0, 1].tolist() == [0, 0, 0, 0]
    assert idx[0, 2].tolist() == [5, 5, 5, 5]
    # more hits than nsample -> first nsample by index
    assert ops.ball_query(new_xyz[:, :1].contiguous(), xyz, 1.0, 2)[0, 0].tolist() == [0, 1]
    # fewer hits than nsample -> padded with the FIRST hit
    assert ops.ball_query(new_xyz[:, :1].contiguous(), xyz, 0.3, 4)[0, 0].tolist() == [0, 3, 0, 0]


def test_three_nn_ties_and_short_known():
    unknown = torch.tensor([[[0., 0, 0]]])
    known = torch.tensor([[[1., 0, 0], [0, 1., 0], [0, 0, 1.], [0.5, 0, 0], [-1., 0, 0]]])
    d2, idx = ops.three_nn(unknown, known)
    assert idx[0, 0].tolist() == [3, 0, 1]          # ties among distance 1 -> lowest indices, in order
    assert torch.allclose(d2[0, 0], torch.tensor([0.25, 1.0, 1.0]))
    d2, idx = ops.three_nn(unknown, known[:, :2].contiguous())
    assert idx[0, 0].tolist() == [0, 1, 0] and torch.isinf(d2[0, 0, 2])   # 1e40 -> inf after the float cast


def test_gather_group_and_grads_accumulate_duplicates():
    feat = torch.arange(2 * 3 * 5, dtype=torch.float32).reshape(2, 3, 5)
    idx = torch.tensor([[4, 4, 0], [1, 1, 1]], dtype=torch.int32)
    out = ops.gather_points(feat, idx)
    assert torch.equal(out, feat.gather(2, idx.long().unsqueeze(1).expand(2, 3, 3)))
    g = ops.gather_points_grad(torch.ones(2, 3, 3), idx, 5)
    assert g[0, :, 4].eq(2).all() and g[0, :, 0].eq(1).all() and g[1, :, 1].eq(3).all() and g.sum() == 18
    gidx = torch.tensor([[[0, 0], [2, 3]], [[4, 4], [4, 1]]], dtype=torch.int32)
    go = ops.group_points(feat, gidx)
    assert go.shape == (2, 3, 2, 2) and go[1, 2, 1, 1] == feat[1, 2, 1] and go[0, 1, 1, 0] == feat[0, 1, 2]
    gg = ops.group_points_grad(torch.ones(2, 3, 2, 2), gidx, 5)
    assert gg[0, 0].tolist() == [2, 0, 1, 1, 0] and gg[1, 0].tolist() == [0, 1, 0, 0, 3]


def test_three_interpolate_and_grad():
    feat = torch.tensor([[[1., 2., 3., 4.]]])
    idx = torch.tensor([[[0, 1, 3], [2, 2, 2]]], dtype=torch.int32)
    w = torch.tensor([[[0.5, 0.25, 0.25], [0.2, 0.3, 0.5]]])
    out = ops.three_interpolate(feat, idx, w)
    assert torch.allclose(out, torch.tensor([[[0.5 + 0.5 + 1.0, 3.0]]]))
    g = ops.three_interpolate_grad(torch.ones(1, 1, 2), idx, w, 4)
    assert torch.allclose(g, torch.tensor([[[0.5, 0.25, 1.0, 0.25]]]))


def test_fps_key_model_matches_tree_emulation():
    """The CUDA kernel (csrc/fps.cu) replaces the shared-memory tree by an arg-max over the total order
    (distance desc, prio asc), prio(k) = bitrev(k mod block) * ceil(N/block) + k div block.  Check that model
    against the literal tree emulation on clouds full of exact duplicates and near-origin points."""
    rng = np.random.default_rng(5)

    def bitrev(t, bits):
        r = 0
        for i in range(bits):
            r |= ((t >> i) & 1) << (bits - 1 - i)
        return r

    for N, npoint in ((37, 20), (64, 40), (100, 60), (700, 64), (1024, 48)):
        base = rng.uniform(-2, 2, size=(max(4, N // 5), 3)).astype(np.float32)
        base[0] = 0.001  # near-origin -> skipped
        pts = base[rng.integers(0, base.shape[0], size=N)]
        want = ops.furthest_point_sampling(torch.from_numpy(pts[None].copy()), npoint)[0].numpy()
        block = ops.opt_n_threads(N)
        bits = int(np.log2(block))
        cnt = (N + block - 1) // block
        prio = np.array([bitrev(k % block, bits) * cnt + k // block for k in range(N)], dtype=np.int64)
        x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
        f32 = np.float32
        mag = np.array([np.float32(np.float64(z[k]) * np.float64(z[k]) + np.float64(
            np.float32(np.float64(y[k]) * np.float64(y[k]) + np.float64(f32(x[k] * x[k]))))) for k in range(N)])
        valid = ~(mag.astype(np.float64) <= 1e-3)
        td = np.full(N, 1e10, dtype=np.float32)
        got, old = [0], 0

        def fma(a, b, c):  # exact for fp32 inputs: the double product/sum is exact before the final rounding
            return np.float32(np.float64(a) * np.float64(b) + np.float64(c))

        for _ in range(1, npoint):
            dx, dy, dz = (x - x[old]).astype(f32), (y - y[old]).astype(f32), (z - z[old]).astype(f32)
            d = np.array([fma(dz[k], dz[k], fma(dy[k], dy[k], f32(dx[k] * dx[k]))) for k in range(N)], dtype=f32)
            td = np.where(valid, np.minimum(d, td), td)
            if not valid.any():
                old = 0
            else:
                cand = np.where(valid)[0]
                best = td[cand].max()
                tied = cand[td[cand] == best]
                old = int(tied[np.argmin(prio[tied])])
            got.append(old)
        assert got == want.tolist(), (N, npoint)
