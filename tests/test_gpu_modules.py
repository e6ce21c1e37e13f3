"""GPU parity of the modules and whole models against (a) the golden vectors produced by the reference's own
Python and (b) the oracle, in both execution modes (fused = product default, composed = op-by-op cross-check)."""
import os

import numpy as np
import pytest
import torch

from _params import det_state_dict
from open3dsot_b200 import runtime
from open3dsot_b200.config import load_config
from open3dsot_b200.datasets.synthetic import synthetic_siamese_batch
from open3dsot_b200.models import get_model
from open3dsot_b200.models.head.rpn import P2BVoteNetRPN
from open3dsot_b200.models.head.xcorr import BoxAwareXCorr, P2B_XCorr
from open3dsot_b200.pointnet2.utils import pointnet2_modules as pm
from open3dsot_b200.pointnet2.utils import pointnet2_utils as pu

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-4
MODES = ["composed"] + (["fused"] if os.path.exists(os.path.join(ROOT, "open3dsot_b200", "fused.py")) else [])


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def gm(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "ref_modules.npz")))


@pytest.fixture(scope="module")
def gmod(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "ref_models.npz")))


@pytest.fixture(params=MODES)
def mode(request):
    old = runtime.fused_enabled()
    runtime.set_fused(request.param == "fused")
    yield request.param
    runtime.set_fused(old)


def _c(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def test_query_and_group_module(gm):
    xyz, feats = _c(gm["qg_xyz"]), _c(gm["qg_feats"])
    new_xyz = xyz[:, :24].contiguous()
    for norm in (0, 1):
        qg = pu.QueryAndGroup(0.35, 16, use_xyz=True, return_idx=True, normalize_xyz=bool(norm))
        nf, idx = qg(xyz, new_xyz, feats)
        assert np.array_equal(idx.cpu().numpy(), gm[f"qg_idx_norm{norm}"])
        assert rel(nf, gm[f"qg_out_norm{norm}"]) < 1e-6


@pytest.mark.parametrize("tag,use_fps", [("fps", True), ("arange", False)])
def test_sa_module(gm, mode, tag, use_fps):
    xyz, feats = _c(gm["qg_xyz"]), _c(gm["qg_feats"])
    sa = pm.PointnetSAModule(mlp=[8, 16, 16, 32], radius=0.35, nsample=16, use_fps=use_fps)
    sa.load_state_dict(det_state_dict(sa.state_dict(), seed=1))
    sa = sa.cuda().train()
    f_in = feats.clone().requires_grad_(True)
    nx, nf, sidx = sa(xyz, f_in, 24, True)
    assert np.array_equal(sidx.cpu().numpy(), gm[f"sa_{tag}_idx"])
    assert rel(nx, gm[f"sa_{tag}_newxyz"]) == 0
    assert rel(nf, gm[f"sa_{tag}_train_out"]) < RTOL
    nf.square().sum().backward()
    assert rel(f_in.grad, gm[f"sa_{tag}_train_gfeat"]) < RTOL
    assert rel(sa.mlps[0].layer0.conv.weight.grad, gm[f"sa_{tag}_train_gw0"]) < RTOL
    assert rel(sa.mlps[0].layer2.conv.weight.grad, gm[f"sa_{tag}_train_gw2"]) < RTOL
    assert rel(sa.mlps[0].layer1.bn.bn.weight.grad, gm[f"sa_{tag}_train_ggamma1"]) < RTOL
    assert rel(sa.mlps[0].layer2.bn.bn.running_mean, gm[f"sa_{tag}_rm2"]) < 1e-5
    assert rel(sa.mlps[0].layer2.bn.bn.running_var, gm[f"sa_{tag}_rv2"]) < 1e-5
    assert int(sa.mlps[0].layer2.bn.bn.num_batches_tracked) == 1
    sa.load_state_dict(det_state_dict(sa.state_dict(), seed=1))
    sa.eval()
    with torch.no_grad():
        _, nf_e, _ = sa(xyz, feats, 24, True)
    assert rel(nf_e, gm[f"sa_{tag}_eval_out"]) < RTOL


def test_fp_module(gm, mode):
    fp = pm.PointnetFPModule(mlp=[12, 16, 12])
    fp.load_state_dict(det_state_dict(fp.state_dict(), seed=2))
    fp = fp.cuda().train()
    kf = _c(gm["fp_kf"]).requires_grad_(True)
    y = fp(_c(gm["fp_unknown"]), _c(gm["fp_known"]), _c(gm["fp_uf"]), kf)
    assert rel(y, gm["fp_out"]) < RTOL
    y.square().sum().backward()
    assert rel(kf.grad, gm["fp_gkf"]) < RTOL


def test_xcorr_and_rpn(gm, mode):
    tf, sf, txyz, sxyz, tbc, sbc = (_c(gm[k]) for k in ("xc_tf", "xc_sf", "xc_txyz", "xc_sxyz", "xc_tbc", "xc_sbc"))
    px = P2B_XCorr(16, 16, 16); px.load_state_dict(det_state_dict(px.state_dict(), seed=3)); px = px.cuda().train()
    assert rel(px(tf, sf, txyz), gm["p2bx_out"]) < RTOL
    bx = BoxAwareXCorr(16, 16, 16, k=4, bc_channel=9); bx.load_state_dict(det_state_dict(bx.state_dict(), seed=4))
    bx = bx.cuda().train()
    assert rel(bx(tf, sf, txyz, sxyz, tbc, sbc), gm["bax_out"]) < RTOL
    rp = P2BVoteNetRPN(16, vote_channel=16, num_proposal=8); rp.load_state_dict(det_state_dict(rp.state_dict(), seed=5))
    rp = rp.cuda().train()
    boxes, cla, vxyz, cen = rp(sxyz, sf)
    assert rel(boxes, gm["rpn_boxes"]) < RTOL and rel(cla, gm["rpn_cla"]) < RTOL
    assert rel(vxyz, gm["rpn_vote_xyz"]) < RTOL and rel(cen, gm["rpn_centers"]) < RTOL


@pytest.mark.parametrize("name,cfg_file,B,seed", [("bat", "BAT_Car.yaml", 2, 21), ("p2b", "P2B_Car.yaml", 1, 22)])
def test_whole_model_against_reference_golden(gmod, mode, name, cfg_file, B, seed):
    cfg = load_config(os.path.join(ROOT, "cfgs", cfg_file))
    net = get_model(cfg.net_model)(cfg)
    base = det_state_dict(net.state_dict(), seed=seed)
    net.load_state_dict(base)
    net = net.cuda().train()
    batch = synthetic_siamese_batch(B, 256, 512, seed=1234 + seed, box_aware=(name == "bat"), device="cuda")
    # Whole-model tolerance is looser than the per-module 1e-4: vote clustering ball-queries COMPUTED coordinates and
    # BoxAware takes a top-k of COMPUTED box clouds, so fp32 round-off differences between the CPU reference run and
    # the GPU can flip a neighbour choice; every module is held to 1e-4 on identical inputs in the tests above.
    with torch.no_grad():
        ep = net(batch)
    assert np.array_equal(ep["sample_idxs"].cpu().numpy(), gmod[f"{name}_sample_idxs"])
    for k in ("estimation_boxes", "estimation_cla", "vote_xyz", "center_xyz"):
        assert rel(ep[k], gmod[f"{name}_{k}"]) < 2e-3, k
    net.load_state_dict(base)
    net.eval()
    with torch.no_grad():
        ep = net(batch)
    assert rel(ep["estimation_boxes"], gmod[f"{name}_eval_boxes"]) < 2e-3
    net.load_state_dict(base)
    net.train()
    loss = net.training_step({k: v.clone() for k, v in batch.items()}, 0)
    assert rel(loss, gmod[f"{name}_loss"]) < 2e-3
    loss.backward()
    sd = dict(net.named_parameters())
    gscale = float(np.max(gmod[f"{name}_gradnorms"]))
    for key in gmod:
        if key.startswith(f"{name}_grad::"):
            p = key.split("::")[1]
            got, want = sd[p].grad[:16].detach().double().cpu(), torch.from_numpy(gmod[key]).double()
            assert float((got - want).norm()) < 2e-2 * float(want.norm()) + 1e-5 * gscale, p
    norms = np.array([float(sd[k].grad.norm()) for k in sorted(sd)])
    assert np.allclose(norms, gmod[f"{name}_gradnorms"], rtol=3e-2, atol=1e-5 * gscale)
