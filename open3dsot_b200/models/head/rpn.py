"""VoteNet-style region proposal head.  Mirror of models/head/rpn.py: P2BVoteNetRPN (:12-67).
Parameter names: FC_layer_cla.{0,1,2}, vote_layer.{0,1,2}, vote_aggregation.mlps.0.layer{0,1,2},
FC_proposal.{0,1,2}."""
import torch
from torch import nn

from ...pointnet2.utils import pytorch_utils as pt_utils
from ...pointnet2.utils.pointnet2_modules import PointnetSAModule
from ... import runtime


class P2BVoteNetRPN(nn.Module):
    def __init__(self, feature_channel, vote_channel=256, num_proposal=64, normalize_xyz=False):
        super().__init__()
        self.num_proposal = num_proposal
        self.FC_layer_cla = (pt_utils.Seq(feature_channel).conv1d(feature_channel, bn=True)
                             .conv1d(feature_channel, bn=True).conv1d(1, activation=None))
        self.vote_layer = (pt_utils.Seq(3 + feature_channel).conv1d(feature_channel, bn=True)
                           .conv1d(feature_channel, bn=True).conv1d(3 + feature_channel, activation=None))
        # vote clustering: the first `num_proposal` votes are the cluster centres (use_fps defaults to False)
        self.vote_aggregation = PointnetSAModule(radius=0.3, nsample=16,
                                                 mlp=[1 + feature_channel, vote_channel, vote_channel, vote_channel],
                                                 use_xyz=True, normalize_xyz=normalize_xyz)
        self.FC_proposal = (pt_utils.Seq(vote_channel).conv1d(vote_channel, bn=True)
                            .conv1d(vote_channel, bn=True).conv1d(3 + 1 + 1, activation=None))

    def forward(self, xyz, feature):
        """xyz (B,N,3), feature (B,f,N) -> boxes (B,num_proposal,5) [x,y,z,theta,objectness],
        seed logits (B,N), vote_xyz (B,N,3), proposal centres (B,num_proposal,3)."""
        if runtime.fused_enabled():
            from ... import fused
            mlp = fused.seq_forward
        else:
            mlp = lambda m, x: m(x)  # noqa: E731
        join_cla = None
        if runtime.fused_enabled() and fused.branch_overlap(feature):
            # inference: the seed classifier and the vote layer read the same features and meet only at the vote clustering
            join_cla = fused.run_ahead(lambda: mlp(self.FC_layer_cla, feature).squeeze(1))
        else:
            estimation_cla = mlp(self.FC_layer_cla, feature).squeeze(1)
        xyz_feature = torch.cat((xyz.transpose(1, 2), feature), dim=1)
        vote = xyz_feature + mlp(self.vote_layer, xyz_feature)
        if join_cla is not None:
            estimation_cla = join_cla()
        score = estimation_cla.sigmoid()
        vote_xyz = vote[:, 0:3, :].transpose(1, 2).contiguous()
        vote_feature = torch.cat((score.unsqueeze(1), vote[:, 3:, :]), dim=1)
        center_xyzs, proposal_features = self.vote_aggregation(vote_xyz, vote_feature, self.num_proposal)
        proposal_offsets = mlp(self.FC_proposal, proposal_features)
        estimation_boxes = torch.cat((proposal_offsets[:, 0:3, :] + center_xyzs.transpose(1, 2),
                                      proposal_offsets[:, 3:5, :]), dim=1).transpose(1, 2).contiguous()
        return estimation_boxes, estimation_cla, vote_xyz, center_xyzs
