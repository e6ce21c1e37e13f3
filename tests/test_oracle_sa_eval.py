"""The float64 composition that tests/test_gpu_sa_fused.py holds the single-kernel inference SA layer to, against the pinned CPU
oracle (oracle/modules.py `sa_module`, eval mode — itself checked against the reference's own Python in test_oracle_golden.py):
same weights, same running statistics, the oracle's ball-query indices.  No GPU."""
import pytest
import torch

from oracle import modules as om
from oracle import ops as oops
from test_gpu_sa_fused import _cloud, _module, _reference64, rel


@pytest.mark.parametrize("C,mlp,N,npoint,radius,S,normalize", [
    (0, [0, 64, 64, 128], 256, 64, 0.3, 32, False),
    (128, [128, 128, 128, 256], 128, 32, 0.5, 32, False),
    (5, [5, 7, 130], 100, 8, 0.4, 16, True),
])
def test_float64_yardstick_equals_the_pinned_oracle(C, mlp, N, npoint, radius, S, normalize):
    B = 2
    xyz, g = _cloud(B, N, seed=11 + C)
    feats = torch.randn(B, C, N, generator=g) * 0.7 if C else None
    sa = _module(mlp, radius, S, seed=3, normalize=normalize, device="cpu")
    sd = {"sa." + k: v.detach().clone() for k, v in sa.state_dict().items()}
    new_xyz = xyz[:, :npoint].contiguous()
    idx = oops.ball_query(new_xyz, xyz, radius, S)
    want, _ = _reference64(sa, xyz, feats, npoint, idx=idx)
    _, got, _ = om.sa_module(sd, "sa", xyz, feats, npoint, radius, S, False, False, normalize_xyz=normalize)
    assert got.shape == want.shape
    assert rel(got, want) < 2e-6                        # the oracle computes in float32
