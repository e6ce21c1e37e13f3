"""Template–search fusion heads.  Mirror of models/head/xcorr.py: BaseXCorr (:10-17), P2B_XCorr (:20-53),
BoxAwareXCorr (:56-103).  Parameter names: `mlp.layer{0,1,2}.*`, `fea_layer.{0,1}.*`; parameter-free child
`cosine` kept for state-dict/module-tree parity."""
import torch
import torch.nn.functional as F
from torch import nn

from ...pointnet2.utils import pointnet2_utils
from ...pointnet2.utils import pytorch_utils as pt_utils
from ... import runtime


class BaseXCorr(nn.Module):
    def __init__(self, in_channel, hidden_channel, out_channel):
        super().__init__()
        self.cosine = nn.CosineSimilarity(dim=1)
        self.mlp = pt_utils.SharedMLP([in_channel, hidden_channel, hidden_channel, hidden_channel], bn=True)
        self.fea_layer = (pt_utils.Seq(hidden_channel).conv1d(hidden_channel, bn=True)
                          .conv1d(out_channel, activation=None))


class P2B_XCorr(BaseXCorr):
    """Point-wise cosine correlation: every search point sees all template points."""

    def __init__(self, feature_channel, hidden_channel, out_channel):
        super().__init__(feature_channel + 4, hidden_channel, out_channel)

    def forward(self, template_feature, search_feature, template_xyz):
        """template_feature (B,f,M), search_feature (B,f,N), template_xyz (B,M,3) -> (B,out,N)."""
        if runtime.fused_enabled() and 128 % template_feature.size(2) == 0:   # group pooling needs n1 | 128
            from ... import fused
            return fused.p2b_xcorr_forward(self, template_feature, search_feature, template_xyz)
        B, f, n1 = template_feature.shape
        n2 = search_feature.size(2)
        t_exp = template_feature.unsqueeze(-1).expand(B, f, n1, n2)
        sim = self.cosine(t_exp, search_feature.unsqueeze(2).expand(B, f, n1, n2))           # (B,n1,n2)
        xyz_exp = template_xyz.transpose(1, 2).contiguous().unsqueeze(-1).expand(B, 3, n1, n2)
        fusion = torch.cat((sim.unsqueeze(1), xyz_exp, t_exp), dim=1)                         # (B,1+3+f,n1,n2)
        fusion = self.mlp(fusion)
        fusion = F.max_pool2d(fusion, kernel_size=[fusion.size(2), 1]).squeeze(2)             # max over template
        return self.fea_layer(fusion)


class BoxAwareXCorr(BaseXCorr):
    """Box-aware correlation: each search point gathers its k nearest template points in box-cloud space."""

    def __init__(self, feature_channel, hidden_channel, out_channel, k=8, use_search_bc=False,
                 use_search_feature=False, bc_channel=9):
        self.k = k
        self.use_search_bc = use_search_bc
        self.use_search_feature = use_search_feature
        mlp_in_channel = feature_channel + 3 + bc_channel
        if use_search_bc:
            mlp_in_channel += bc_channel
        if use_search_feature:
            mlp_in_channel += feature_channel
        super().__init__(mlp_in_channel, hidden_channel, out_channel)

    def forward(self, template_feature, search_feature, template_xyz, search_xyz=None, template_bc=None,
                search_bc=None):
        """template_feature (B,f,M), search_feature (B,f,N), template_xyz (B,M,3), template_bc (B,M,9),
        search_bc (B,N,9) -> (B,out,N)."""
        if runtime.fused_enabled() and 128 % self.k == 0 and not (self.use_search_bc or self.use_search_feature):
            from ... import fused
            return fused.boxaware_xcorr_forward(self, template_feature, search_feature, template_xyz, template_bc,
                                                search_bc)
        dist_matrix = torch.cdist(template_bc, search_bc)                                      # (B,M,N)
        tmpl = torch.cat([template_xyz.transpose(1, 2), template_bc.transpose(1, 2), template_feature], dim=1)
        topk = torch.argsort(dist_matrix, dim=1, stable=True)[:, :self.k, :].transpose(1, 2).contiguous().int()
        corr = pointnet2_utils.grouping_operation(tmpl.contiguous(), topk)                     # (B,3+9+f,N,k)
        if self.use_search_bc:
            corr = torch.cat([search_bc.transpose(1, 2).unsqueeze(-1).expand(-1, -1, -1, self.k), corr], dim=1)
        if self.use_search_feature:
            corr = torch.cat([search_feature.unsqueeze(-1).expand(-1, -1, -1, self.k), corr], dim=1)
        fusion = self.mlp(corr).max(dim=-1)[0]
        return self.fea_layer(fusion)
