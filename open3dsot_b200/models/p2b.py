"""P2B.  Mirror of models/p2b.py: __init__ (:14-26), forward (:28-59), training_step (:61-101)."""
from torch import nn

from . import base_model
from .backbone.pointnet import Pointnet_Backbone
from .head.rpn import P2BVoteNetRPN
from .head.xcorr import P2B_XCorr
from .. import runtime


class P2B(base_model.MatchingBaseModel):
    def __init__(self, config=None, **kwargs):
        super().__init__(config, **kwargs)
        self.save_hyperparameters()
        c = self.config
        self.backbone = Pointnet_Backbone(c.use_fps, c.normalize_xyz, return_intermediate=False)
        self.conv_final = nn.Conv1d(256, c.feature_channel, kernel_size=1)
        self.xcorr = P2B_XCorr(feature_channel=c.feature_channel, hidden_channel=c.hidden_channel,
                               out_channel=c.out_channel)
        self.rpn = P2BVoteNetRPN(c.feature_channel, vote_channel=c.vote_channel, num_proposal=c.num_proposal,
                                 normalize_xyz=c.normalize_xyz)

    def _pointwise(self, module, x):
        if runtime.fused_enabled():
            from .. import fused
            return fused.seq_forward(module, x)
        return module(x)

    def forward(self, input_dict):
        """input_dict: template_points (B,M,3), search_points (B,N,3) [+ labels]."""
        template, search = input_dict['template_points'], input_dict['search_points']
        M, N = template.shape[1], search.shape[1]
        fused = None
        if runtime.fused_enabled() and search.is_cuda:
            from .. import fused

        def template_branch():
            xyz, feat, _ = self.backbone(template, [M // 2, M // 4, M // 8])
            return xyz, self._pointwise(self.conv_final, feat)

        if fused is not None and fused.branch_overlap(search):
            join_t = fused.run_ahead(template_branch)        # inference: template branch on the side stream
            search_xyz, search_feature, sample_idxs = self.backbone(search, [N // 2, N // 4, N // 8])
            search_feature = self._pointwise(self.conv_final, search_feature)
            template_xyz, template_feature = join_t()
        else:
            join = None
            if self.config.use_fps and fused is not None:
                join = fused.fps_ahead(search, N // 2)
            template_xyz, template_feature = template_branch()
            search_xyz, search_feature, sample_idxs = self.backbone(search, [N // 2, N // 4, N // 8],
                                                                    first_sample_idxs=join() if join else None)
            search_feature = self._pointwise(self.conv_final, search_feature)
        fusion_feature = self.xcorr(template_feature, search_feature, template_xyz)
        estimation_boxes, estimation_cla, vote_xyz, center_xyzs = self.rpn(search_xyz, fusion_feature)
        return {"estimation_boxes": estimation_boxes, "vote_center": vote_xyz, "pred_seg_score": estimation_cla,
                "center_xyz": center_xyzs, 'sample_idxs': sample_idxs, 'estimation_cla': estimation_cla,
                "vote_xyz": vote_xyz}

    def training_step(self, batch, batch_idx):
        end_points = self(batch)
        N = end_points['estimation_cla'].shape[1]
        batch["seg_label"] = batch['seg_label'].gather(dim=1, index=end_points['sample_idxs'][:, :N].long())
        loss_dict = self.compute_loss(batch, end_points)
        c = self.config
        loss = (loss_dict['loss_objective'] * c.objectiveness_weight + loss_dict['loss_box'] * c.box_weight
                + loss_dict['loss_seg'] * c.seg_weight + loss_dict['loss_vote'] * c.vote_weight)
        self.log('loss/train', loss.detach(), on_step=True, on_epoch=True, prog_bar=True, logger=False)
        for k, v in loss_dict.items():
            self.log(f'{k}/train', v.detach(), on_step=True, on_epoch=True, prog_bar=True, logger=False)
        return loss
