"""ORACLE (test infrastructure): functional torch-CPU restatement of the reference composition.

Everything here is a *function of a state dict* that uses the reference's parameter names
(SURVEY.md §8b), so the same weights can be fed to the product modules and to the oracle.
Each function cites the reference lines it restates (paths relative to /root/reference).

Index ops come from the C oracle (oracle/ops.py); the differentiable gathers are expressed
with torch indexing so CPU autograd supplies the oracle gradients (deterministic order).
"""
import torch
import torch.nn.functional as F

from . import ops

BN_EPS = 1e-5
BN_MOMENTUM = 0.1

# Optional recorder of intermediates (tests): set to a dict; every call appends to TAPS[name] in call order
# (the shared backbone runs on the template first, then on the search area: bat.py:89-90).
TAPS = None


# Optional substitution of the DISCRETE choices (ball-query indices, box-cloud top-k), by kind and call order: the parity tests
# run the oracle a second time in float64 (the "exact" arithmetic) with the float32 run's choices, so that the two runs differ by
# floating-point round-off only.  FORCE = {"ball_query": [idx, ...], "topk": [idx]}; None = compute.
FORCE = None
_calls = {}


def _choice(kind, compute):
    if FORCE is None:
        return compute()
    n = _calls.get(kind, 0)
    _calls[kind] = n + 1
    forced = FORCE.get(kind, [])
    return forced[n] if n < len(forced) and forced[n] is not None else compute()


def set_force(force):
    global FORCE
    FORCE = force
    _calls.clear()


def _tap(name, value):
    if TAPS is not None:
        TAPS.setdefault(name, []).append(value.detach().clone() if torch.is_tensor(value) else value)


# --------------------------------------------------------------------------- index helpers
def gather_cols(features, idx):
    """(B,C,N),(B,M)i32 -> (B,C,M); differentiable form of `_ext.gather_points`
    (pointnet2/utils/pointnet2_utils.py:68-102)."""
    B, C, _ = features.shape
    return features.gather(2, idx.long().unsqueeze(1).expand(B, C, idx.shape[1]))


def group_cols(features, idx):
    """(B,C,N),(B,M,S)i32 -> (B,C,M,S); differentiable form of `_ext.group_points`
    (pointnet2/utils/pointnet2_utils.py:194-242)."""
    B, C, _ = features.shape
    _, M, S = idx.shape
    flat = idx.long().reshape(B, 1, M * S).expand(B, C, M * S)
    return features.gather(2, flat).reshape(B, C, M, S)


# --------------------------------------------------------------------------- layer builders
def _bn(x, sd, prefix, training):
    """_BNBase (pointnet2/utils/pytorch_utils.py:40-47): child named `bn` -> keys `<prefix>.bn.bn.*`.
    Train mode: biased batch variance for normalisation, running stats updated in `sd` with
    unbiased variance and momentum 0.1 (torch semantics)."""
    w, b = sd[prefix + ".bn.bn.weight"], sd[prefix + ".bn.bn.bias"]
    rm, rv = sd[prefix + ".bn.bn.running_mean"], sd[prefix + ".bn.bn.running_var"]
    y = F.batch_norm(x, rm, rv, w, b, training=training, momentum=BN_MOMENTUM, eps=BN_EPS)
    if training and (prefix + ".bn.bn.num_batches_tracked") in sd:
        sd[prefix + ".bn.bn.num_batches_tracked"] += 1
    return y


def conv_unit(x, sd, prefix, training, relu=True):
    """One `_ConvBase` (pytorch_utils.py:68-121): 1x1 conv (bias only when no BN) -> BN -> ReLU."""
    w = sd[prefix + ".conv.weight"]
    bias = sd.get(prefix + ".conv.bias")
    has_bn = (prefix + ".bn.bn.weight") in sd
    if w.dim() == 4:
        y = F.conv2d(x, w, bias)
    else:
        y = F.conv1d(x, w, bias)
    if has_bn:
        y = _bn(y, sd, prefix, training)
    if relu:
        y = F.relu(y)
    return y


def shared_mlp(x, sd, prefix, training):
    """SharedMLP (pytorch_utils.py:12-37): children `layer0..layerK`, all conv+BN+ReLU."""
    i = 0
    while (f"{prefix}.layer{i}.conv.weight") in sd:
        x = conv_unit(x, sd, f"{prefix}.layer{i}", training, relu=True)
        i += 1
    return x


def seq_conv1d(x, sd, prefix, training, last_has_activation=False):
    """`pt_utils.Seq(...).conv1d(...)...` (pytorch_utils.py:300-339): children "0","1",...;
    every use in the hot path ends with `activation=None` on the last conv
    (models/head/rpn.py:17-21, models/bat.py:22-25, models/head/xcorr.py:15-17)."""
    n = 0
    while (f"{prefix}.{n}.conv.weight") in sd:
        n += 1
    for i in range(n):
        last = i == n - 1
        x = conv_unit(x, sd, f"{prefix}.{i}", training, relu=(not last) or last_has_activation)
    return x


# --------------------------------------------------------------------------- PointNet++ layers
def query_and_group(xyz, new_xyz, features, radius, nsample, use_xyz=True, normalize_xyz=False):
    """QueryAndGroup.forward (pointnet2/utils/pointnet2_utils.py:299-339)."""
    idx = _choice("ball_query", lambda: ops.ball_query(new_xyz.detach().float().contiguous(), xyz.detach().float().contiguous(),
                                                      radius, nsample))
    grouped_xyz = group_cols(xyz.transpose(1, 2), idx) - new_xyz.transpose(1, 2).unsqueeze(-1)
    if normalize_xyz:
        grouped_xyz = grouped_xyz / radius
    if features is None:
        assert use_xyz
        return grouped_xyz, idx
    grouped = group_cols(features, idx)
    if use_xyz:
        grouped = torch.cat([grouped_xyz, grouped], dim=1)  # channel order [xyz(3), C]
    return grouped, idx


def sa_module(sd, prefix, xyz, features, npoint, radius, nsample, use_fps, training,
              normalize_xyz=False, use_xyz=True):
    """_PointnetSAModuleBase.forward with a single scale (pointnet2/utils/pointnet2_modules.py:31-79):
    FPS or the first `npoint` points as centres -> group -> SharedMLP -> max over nsample."""
    B = xyz.shape[0]
    if use_fps:
        sample_idxs = ops.furthest_point_sampling(xyz.detach().float().contiguous(), npoint)
    else:
        sample_idxs = torch.arange(npoint, dtype=torch.int32).repeat(B, 1)
    new_xyz = gather_cols(xyz.transpose(1, 2), sample_idxs).transpose(1, 2).contiguous()
    grouped, bq_idx = query_and_group(xyz, new_xyz, features, radius, nsample, use_xyz, normalize_xyz)
    y = shared_mlp(grouped, sd, f"{prefix}.mlps.0", training)
    y = F.max_pool2d(y, kernel_size=[1, y.size(3)]).squeeze(-1)
    _tap(f"{prefix}:bq_idx", bq_idx)
    _tap(f"{prefix}:out", y)
    return new_xyz, y, sample_idxs


def fp_module(sd, prefix, unknown, known, unknow_feats, known_feats, training):
    """PointnetFPModule.forward (pointnet2/utils/pointnet2_modules.py:168-212)."""
    dist2, idx = ops.three_nn(unknown.detach().contiguous(), known.detach().contiguous())
    dist = torch.sqrt(dist2)  # ThreeNN.forward, pointnet2_utils.py:127
    dist_recip = 1.0 / (dist + 1e-8)
    weight = dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)
    B, c, m = known_feats.shape
    n = unknown.shape[1]
    g = group_cols(known_feats, idx)  # (B,c,n,3)
    interpolated = (g * weight.unsqueeze(1)).sum(-1)
    new = interpolated if unknow_feats is None else torch.cat([interpolated, unknow_feats], dim=1)
    return shared_mlp(new.unsqueeze(-1), sd, f"{prefix}.mlp", training).squeeze(-1)


SA_SPECS = [(0.3, 32), (0.5, 32), (0.7, 32)]  # models/backbone/pointnet.py:32-58


def backbone(sd, prefix, pc, numpoints, use_fps, training, normalize_xyz=False):
    """Pointnet_Backbone.forward (models/backbone/pointnet.py:66-88); FPS only in SA1."""
    xyz = pc[..., 0:3].contiguous()
    feats = pc[..., 3:].transpose(1, 2).contiguous() if pc.size(-1) > 3 else None
    idx0 = None
    for i, (r, ns) in enumerate(SA_SPECS):
        xyz, feats, sidx = sa_module(sd, f"{prefix}.SA_modules.{i}", xyz, feats, numpoints[i], r, ns,
                                     use_fps and i == 0, training, normalize_xyz)
        if i == 0:
            idx0 = sidx
    return xyz, feats, idx0


# --------------------------------------------------------------------------- fusion heads
def boxaware_xcorr(sd, prefix, t_feat, s_feat, t_xyz, s_xyz, t_bc, s_bc, k, training):
    """BoxAwareXCorr.forward (models/head/xcorr.py:67-103) with use_search_bc/use_search_feature False
    (the only setting any shipped cfg uses; the other branches reference an undefined `self.K`)."""
    dist = torch.cdist(t_bc, s_bc)  # (B,M,N), matmul formulation for these sizes
    topk = _choice("topk", lambda: torch.argsort(dist, dim=1, stable=True)[:, :k, :]   # reference: unstable argsort (ties undefined)
                   .transpose(1, 2).contiguous().int())  # (B,N,k)
    tmpl = torch.cat([t_xyz.transpose(1, 2), t_bc.transpose(1, 2), t_feat], dim=1)
    corr = group_cols(tmpl, topk)  # (B,3+9+D,N,k)
    y = shared_mlp(corr, sd, f"{prefix}.mlp", training)
    y = y.max(dim=-1)[0]
    out = seq_conv1d(y, sd, f"{prefix}.fea_layer", training)
    _tap(f"{prefix}:topk", topk)
    _tap(f"{prefix}:out", out)
    return out, topk


def p2b_xcorr(sd, prefix, t_feat, s_feat, t_xyz, training):
    """P2B_XCorr.forward (models/head/xcorr.py:25-53)."""
    B, f, n1 = t_feat.shape
    n2 = s_feat.shape[2]
    sim = F.cosine_similarity(t_feat.unsqueeze(-1).expand(B, f, n1, n2),
                              s_feat.unsqueeze(2).expand(B, f, n1, n2), dim=1)  # eps 1e-8
    fusion = torch.cat([sim.unsqueeze(1),
                        t_xyz.transpose(1, 2).unsqueeze(-1).expand(B, 3, n1, n2),
                        t_feat.unsqueeze(-1).expand(B, f, n1, n2)], dim=1)
    y = shared_mlp(fusion, sd, f"{prefix}.mlp", training)
    y = F.max_pool2d(y, kernel_size=[y.size(2), 1]).squeeze(2)
    out = seq_conv1d(y, sd, f"{prefix}.fea_layer", training)
    _tap(f"{prefix}:out", out)
    return out


def rpn(sd, prefix, xyz, feature, num_proposal, training, normalize_xyz=False):
    """P2BVoteNetRPN.forward (models/head/rpn.py:41-67)."""
    cla = seq_conv1d(feature, sd, f"{prefix}.FC_layer_cla", training).squeeze(1)
    score = cla.sigmoid()
    xyz_feature = torch.cat((xyz.transpose(1, 2), feature), dim=1)
    vote = xyz_feature + seq_conv1d(xyz_feature, sd, f"{prefix}.vote_layer", training)
    vote_xyz = vote[:, 0:3, :].transpose(1, 2).contiguous()
    vote_feature = torch.cat((score.unsqueeze(1), vote[:, 3:, :]), dim=1)
    centers, prop_feat, _ = sa_module(sd, f"{prefix}.vote_aggregation", vote_xyz, vote_feature, num_proposal,
                                      0.3, 16, False, training, normalize_xyz)
    offs = seq_conv1d(prop_feat, sd, f"{prefix}.FC_proposal", training)
    boxes = torch.cat((offs[:, 0:3, :] + centers.transpose(1, 2), offs[:, 3:5, :]), dim=1)
    return boxes.transpose(1, 2).contiguous(), cla, vote_xyz, centers


# --------------------------------------------------------------------------- whole models
def bat_forward(sd, cfg, batch, training):
    """BAT.forward (models/bat.py:67-112)."""
    template, search, template_bc = batch["template_points"], batch["search_points"], batch["points2cc_dist_t"]
    M, N = template.shape[1], search.shape[1]
    t_xyz, t_feat, idx_t = backbone(sd, "backbone", template, [M // 2, M // 4, M // 8], cfg["use_fps"], training,
                                    cfg["normalize_xyz"])
    s_xyz, s_feat, idx_s = backbone(sd, "backbone", search, [N // 2, N // 4, N // 8], cfg["use_fps"], training,
                                    cfg["normalize_xyz"])
    t_feat = F.conv1d(t_feat, sd["conv_final.weight"], sd["conv_final.bias"])
    s_feat = F.conv1d(s_feat, sd["conv_final.weight"], sd["conv_final.bias"])
    pred_bc = seq_conv1d(torch.cat([s_xyz.transpose(1, 2), s_feat], dim=1), sd, "mlp_bc", training).transpose(1, 2)
    sel = idx_t[:, :M // 8, None].repeat(1, 1, cfg["bc_channel"]).long()
    t_bc = template_bc.gather(dim=1, index=sel)
    fusion, topk = boxaware_xcorr(sd, "xcorr", t_feat, s_feat, t_xyz, s_xyz, t_bc, pred_bc, cfg["k"], training)
    boxes, cla, vote_xyz, centers = rpn(sd, "rpn", s_xyz, fusion, cfg["num_proposal"], training, cfg["normalize_xyz"])
    return {"estimation_boxes": boxes, "vote_center": vote_xyz, "pred_seg_score": cla, "center_xyz": centers,
            "sample_idxs": idx_s, "estimation_cla": cla, "vote_xyz": vote_xyz, "pred_search_bc": pred_bc,
            "_xcorr_topk": topk}


def p2b_forward(sd, cfg, batch, training):
    """P2B.forward (models/p2b.py:28-59)."""
    template, search = batch["template_points"], batch["search_points"]
    M, N = template.shape[1], search.shape[1]
    t_xyz, t_feat, _ = backbone(sd, "backbone", template, [M // 2, M // 4, M // 8], cfg["use_fps"], training,
                                cfg["normalize_xyz"])
    s_xyz, s_feat, idx_s = backbone(sd, "backbone", search, [N // 2, N // 4, N // 8], cfg["use_fps"], training,
                                    cfg["normalize_xyz"])
    t_feat = F.conv1d(t_feat, sd["conv_final.weight"], sd["conv_final.bias"])
    s_feat = F.conv1d(s_feat, sd["conv_final.weight"], sd["conv_final.bias"])
    fusion = p2b_xcorr(sd, "xcorr", t_feat, s_feat, t_xyz, training)
    boxes, cla, vote_xyz, centers = rpn(sd, "rpn", s_xyz, fusion, cfg["num_proposal"], training, cfg["normalize_xyz"])
    return {"estimation_boxes": boxes, "vote_center": vote_xyz, "pred_seg_score": cla, "center_xyz": centers,
            "sample_idxs": idx_s, "estimation_cla": cla, "vote_xyz": vote_xyz}


def matching_loss(data, output):
    """MatchingBaseModel.compute_loss (models/base_model.py:122-164)."""
    boxes, cla = output["estimation_boxes"], output["estimation_cla"]
    seg_label, box_label = data["seg_label"], data["box_label"]
    centers, vote_xyz = output["center_xyz"], output["vote_xyz"]
    loss_seg = F.binary_cross_entropy_with_logits(cla, seg_label)
    loss_vote = F.smooth_l1_loss(vote_xyz, box_label[:, None, :3].expand_as(vote_xyz), reduction="none")
    loss_vote = (loss_vote.mean(2) * seg_label).sum() / (seg_label.sum() + 1e-06)
    dist = torch.sqrt(torch.sum((centers - box_label[:, None, :3]) ** 2, dim=-1) + 1e-6)
    obj_label = (dist < 0.3).float()
    obj_mask = ((dist < 0.3) | (dist > 0.6)).float()
    # default MEAN reduction, as the reference calls it (base_model.py:151): the mask multiplies the scalar mean
    loss_obj = F.binary_cross_entropy_with_logits(boxes[:, :, 4], obj_label, pos_weight=torch.tensor([2.0]))
    loss_obj = torch.sum(loss_obj * obj_mask) / (torch.sum(obj_mask) + 1e-6)
    loss_box = F.smooth_l1_loss(boxes[:, :, :4], box_label[:, None, :4].expand_as(boxes[:, :, :4]), reduction="none")
    loss_box = torch.sum(loss_box.mean(2) * obj_label) / (obj_label.sum() + 1e-6)
    return {"loss_objective": loss_obj, "loss_box": loss_box, "loss_seg": loss_seg, "loss_vote": loss_vote}


def bat_training_loss(sd, cfg, batch):
    """BAT.training_step minus logging (models/bat.py:114-145) + BAT.compute_loss (:57-65)."""
    out = bat_forward(sd, cfg, batch, training=True)
    n = out["estimation_cla"].shape[1]
    sidx = out["sample_idxs"][:, :n].long()
    data = dict(batch)
    data["seg_label"] = batch["seg_label"].gather(1, sidx)
    data["points2cc_dist_s"] = batch["points2cc_dist_s"].gather(
        1, sidx[:, :, None].repeat(1, 1, cfg["bc_channel"]))
    ld = matching_loss(data, out)
    loss_bc = F.smooth_l1_loss(out["pred_search_bc"], data["points2cc_dist_s"], reduction="none")
    ld["loss_bc"] = torch.sum(loss_bc.mean(2) * data["seg_label"]) / (data["seg_label"].sum() + 1e-6)
    loss = (ld["loss_objective"] * cfg["objectiveness_weight"] + ld["loss_box"] * cfg["box_weight"]
            + ld["loss_seg"] * cfg["seg_weight"] + ld["loss_vote"] * cfg["vote_weight"]
            + ld["loss_bc"] * cfg["bc_weight"])
    return loss, ld, out


def p2b_training_loss(sd, cfg, batch):
    """P2B.training_step minus logging (models/p2b.py:61-84)."""
    out = p2b_forward(sd, cfg, batch, training=True)
    n = out["estimation_cla"].shape[1]
    data = dict(batch)
    data["seg_label"] = batch["seg_label"].gather(1, out["sample_idxs"][:, :n].long())
    ld = matching_loss(data, out)
    loss = (ld["loss_objective"] * cfg["objectiveness_weight"] + ld["loss_box"] * cfg["box_weight"]
            + ld["loss_seg"] * cfg["seg_weight"] + ld["loss_vote"] * cfg["vote_weight"])
    return loss, ld, out
