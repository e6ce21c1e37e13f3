"""BAT (Box-Aware Tracker).  Mirror of models/bat.py: __init__ (:17-41), compute_loss (:57-65),
forward (:67-112), training_step (:114-166).  Logging is asynchronous: the reference issues twelve
`.item()` host syncs per step (bat.py:146-163); here the loss dict stays on the device and `self.log` receives
detached tensors."""
import torch
import torch.nn.functional as F
from torch import nn

from . import base_model
from .backbone.pointnet import Pointnet_Backbone
from .head.rpn import P2BVoteNetRPN
from .head.xcorr import BoxAwareXCorr
from ..pointnet2.utils import pytorch_utils as pt_utils
from .. import runtime


class BAT(base_model.MatchingBaseModel):
    def __init__(self, config=None, **kwargs):
        super().__init__(config, **kwargs)
        self.save_hyperparameters()
        c = self.config
        self.backbone = Pointnet_Backbone(c.use_fps, c.normalize_xyz, return_intermediate=False)
        self.conv_final = nn.Conv1d(256, c.feature_channel, kernel_size=1)
        self.mlp_bc = (pt_utils.Seq(3 + c.feature_channel).conv1d(c.feature_channel, bn=True)
                       .conv1d(c.feature_channel, bn=True).conv1d(c.bc_channel, activation=None))
        self.xcorr = BoxAwareXCorr(feature_channel=c.feature_channel, hidden_channel=c.hidden_channel,
                                   out_channel=c.out_channel, k=c.k, use_search_bc=c.use_search_bc,
                                   use_search_feature=c.use_search_feature, bc_channel=c.bc_channel)
        self.rpn = P2BVoteNetRPN(c.feature_channel, vote_channel=c.vote_channel, num_proposal=c.num_proposal,
                                 normalize_xyz=c.normalize_xyz)

    def prepare_input(self, template_pc, search_pc, template_box):
        """bat.py:41-55: the matching models' input plus the template's BoxCloud (distances to centre + 8 corners)."""
        from ..tracking import boxes as bx
        from .base_model import regularize
        template_points, _ = regularize(template_pc, self.config.template_size, seed=1)
        search_points, _ = regularize(search_pc, self.config.search_size, seed=1)
        template_bc = bx.point_to_box_distance(template_points, template_box)
        return {'template_points': template_points[None], 'search_points': search_points[None],
                'points2cc_dist_t': template_bc[None]}

    def compute_loss(self, data, output):
        out_dict = super().compute_loss(data, output)
        seg_label = data['seg_label']
        loss_bc = F.smooth_l1_loss(output['pred_search_bc'], data['points2cc_dist_s'], reduction='none')
        out_dict["loss_bc"] = torch.sum(loss_bc.mean(2) * seg_label) / (seg_label.sum() + 1e-6)
        return out_dict

    def _pointwise(self, module, x):
        if runtime.fused_enabled():
            from .. import fused
            return fused.seq_forward(module, x)
        return module(x)

    def forward(self, input_dict):
        """input_dict: template_points (B,M,3), search_points (B,N,3), points2cc_dist_t (B,M,9) [+ labels]."""
        template, search = input_dict['template_points'], input_dict['search_points']
        bc_t = input_dict['points2cc_dist_t']
        M, N = template.shape[1], search.shape[1]
        fused = None
        if runtime.fused_enabled() and search.is_cuda:
            from .. import fused

        def template_branch():
            xyz, feat, idxs = self.backbone(template, [M // 2, M // 4, M // 8])
            sel = idxs[:, :M // 8, None].expand(-1, -1, self.config.bc_channel).long()
            return xyz, self._pointwise(self.conv_final, feat), bc_t.gather(dim=1, index=sel)

        if fused is not None and fused.branch_overlap(search):
            # inference: the two branches are independent up to the cross-correlation -> two streams (two graph branches)
            join_t = fused.run_ahead(template_branch)
            search_xyz, search_feature, sample_idxs = self.backbone(search, [N // 2, N // 4, N // 8])
            search_feature = self._pointwise(self.conv_final, search_feature)
            template_xyz, template_feature, template_bc = join_t()
        else:
            join = None
            if self.config.use_fps and fused is not None:
                join = fused.fps_ahead(search, N // 2)           # search-branch FPS runs underneath the template branch
            template_xyz, template_feature, template_bc = template_branch()
            search_xyz, search_feature, sample_idxs = self.backbone(search, [N // 2, N // 4, N // 8],
                                                                    first_sample_idxs=join() if join else None)
            search_feature = self._pointwise(self.conv_final, search_feature)
        pred_search_bc = self._pointwise(self.mlp_bc, torch.cat([search_xyz.transpose(1, 2), search_feature], dim=1))
        pred_search_bc = pred_search_bc.transpose(1, 2)                                   # (B, N//8, 9)
        fusion_feature = self.xcorr(template_feature, search_feature, template_xyz, search_xyz, template_bc,
                                    pred_search_bc)
        estimation_boxes, estimation_cla, vote_xyz, center_xyzs = self.rpn(search_xyz, fusion_feature)
        return {"estimation_boxes": estimation_boxes, "vote_center": vote_xyz, "pred_seg_score": estimation_cla,
                "center_xyz": center_xyzs, 'sample_idxs': sample_idxs, 'estimation_cla': estimation_cla,
                "vote_xyz": vote_xyz, "pred_search_bc": pred_search_bc}

    def training_step(self, batch, batch_idx):
        end_points = self(batch)
        N = end_points['estimation_cla'].shape[1]
        sidx = end_points['sample_idxs'][:, :N].long()
        batch['seg_label'] = batch['seg_label'].gather(dim=1, index=sidx)
        batch['points2cc_dist_s'] = batch['points2cc_dist_s'].gather(
            dim=1, index=sidx[:, :, None].expand(-1, -1, self.config.bc_channel))
        loss_dict = self.compute_loss(batch, end_points)
        c = self.config
        loss = (loss_dict['loss_objective'] * c.objectiveness_weight + loss_dict['loss_box'] * c.box_weight
                + loss_dict['loss_seg'] * c.seg_weight + loss_dict['loss_vote'] * c.vote_weight
                + loss_dict['loss_bc'] * c.bc_weight)
        self.log('loss/train', loss.detach(), on_step=True, on_epoch=True, prog_bar=True, logger=False)
        for k, v in loss_dict.items():
            self.log(f'{k}/train', v.detach(), on_step=True, on_epoch=True, prog_bar=True, logger=False)
        return loss
