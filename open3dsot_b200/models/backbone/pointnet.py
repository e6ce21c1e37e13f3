"""Backbones.  Mirror of models/backbone/pointnet.py: Pointnet_Backbone (:12-88) — three single-scale SA
layers (r = 0.3/0.5/0.7, nsample 32, MLPs [C,64,64,128], [128,128,128,256], [256,256,256,256]; FPS only in
SA1) — plus the dense M2-Track nets MiniPointNet (:91-141) and SegPointNet (:144-204)."""
import torch
import torch.nn as nn

from ...pointnet2.utils.pointnet2_modules import PointnetSAModule
from ... import runtime

_SA_SPECS = ((0.3, (64, 64, 128)), (0.5, (128, 128, 256)), (0.7, (256, 256, 256)))


class Pointnet_Backbone(nn.Module):
    def __init__(self, use_fps=False, normalize_xyz=False, return_intermediate=False, input_channels=0):
        super().__init__()
        self.return_intermediate = return_intermediate
        self.SA_modules = nn.ModuleList()
        c_in = input_channels
        for i, (radius, widths) in enumerate(_SA_SPECS):
            self.SA_modules.append(PointnetSAModule(radius=radius, nsample=32, mlp=[c_in, *widths], use_xyz=True,
                                                    use_fps=use_fps and i == 0, normalize_xyz=normalize_xyz))
            c_in = widths[-1]

    def _break_up_pc(self, pc):
        xyz = pc[..., 0:3].contiguous()
        features = pc[..., 3:].transpose(1, 2).contiguous() if pc.size(-1) > 3 else None
        return xyz, features

    def forward(self, pointcloud, numpoints, first_sample_idxs=None):
        """pointcloud (B,N,3+C), numpoints [n1,n2,n3] -> xyz (B,n3,3), features (B,256,n3), SA1 sample idx (B,n1).
        `first_sample_idxs` (extension): SA1's FPS indices when they were computed ahead of time."""
        xyz, features = self._break_up_pc(pointcloud)
        l_xyz, l_features, l_idxs = [xyz], [features], []
        for i, sa in enumerate(self.SA_modules):
            li_xyz, li_features, sample_idxs = sa(l_xyz[i], l_features[i], numpoints[i], True,
                                                  sample_idxs=first_sample_idxs if i == 0 else None)
            l_xyz.append(li_xyz)
            l_features.append(li_features)
            l_idxs.append(sample_idxs)
        if self.return_intermediate:
            return l_xyz[1:], l_features[1:], l_idxs[0]
        return l_xyz[-1], l_features[-1], l_idxs[0]


def _conv_bn_relu(c_in, c_out):
    return nn.Sequential(nn.Conv1d(c_in, c_out, 1), nn.BatchNorm1d(c_out), nn.ReLU())


class MiniPointNet(nn.Module):
    """Per-point Conv1d+BN+ReLU stack, global max-pool, then a small FC head (M2-Track stage 2).
    State-dict keys: `features.{i}.*` (one flat Sequential, as in the reference) and `fc.*`."""

    def __init__(self, input_channel, per_point_mlp, hidden_mlp, output_size=0):
        super().__init__()
        layers, c = [], input_channel
        for w in per_point_mlp:
            layers += [nn.Conv1d(c, w, 1), nn.BatchNorm1d(w), nn.ReLU()]
            c = w
        layers += [nn.AdaptiveMaxPool1d(output_size=1), nn.Flatten()]
        for w in hidden_mlp:
            layers += [nn.Linear(c, w), nn.BatchNorm1d(w), nn.ReLU()]
            c = w
        self.features = nn.Sequential(*layers)
        self.output_size = output_size
        if output_size >= 0:
            self.fc = nn.Linear(c, output_size)

    def forward(self, x):
        """x (B,C,N) -> (B,output_size)."""
        if runtime.fused_enabled() and x.is_cuda:
            from ... import fused
            return fused.minipointnet_forward(self, x)
        x = self.features(x)
        return self.fc(x) if self.output_size > 0 else x


class SegPointNet(nn.Module):
    """Per-point segmentation net: per-point features ⊕ broadcast global max feature -> per-point logits."""

    def __init__(self, input_channel, per_point_mlp1, per_point_mlp2, output_size=0, return_intermediate=False):
        super().__init__()
        self.return_intermediate = return_intermediate
        self.seq_per_point = nn.ModuleList()
        c = input_channel
        for w in per_point_mlp1:
            self.seq_per_point.append(_conv_bn_relu(c, w))
            c = w
        self.pool = nn.AdaptiveMaxPool1d(output_size=1)
        self.seq_per_point2 = nn.ModuleList()
        c = c + per_point_mlp1[1]
        for w in per_point_mlp2:
            self.seq_per_point2.append(_conv_bn_relu(c, w))
            c = w
        self.output_size = output_size
        if output_size >= 0:
            self.fc = nn.Conv1d(c, output_size, 1)

    def forward(self, x):
        """x (B,C,N) -> (B,output_size,N) [, intermediate features]."""
        if runtime.fused_enabled() and x.is_cuda:
            from ... import fused
            return fused.segpointnet_forward(self, x)
        second = None
        for i, layer in enumerate(self.seq_per_point):
            x = layer(x)
            if i == 1:
                second = x
        pooled_feature = self.pool(x)  # (B,C,1)
        x = torch.cat([second, pooled_feature.expand_as(x)], dim=1)
        for layer in self.seq_per_point2:
            x = layer(x)
        if self.output_size > 0:
            x = self.fc(x)
        if self.return_intermediate:
            return x, pooled_feature.squeeze(dim=-1)
        return x
