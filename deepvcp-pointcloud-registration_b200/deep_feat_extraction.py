"""feat_extraction_layer -- deep_feat_extraction.py:5-32 of the reference.

The reference's forward crashes after sa1 at HEAD (it feeds normals into layers
declared for 32/64 channels); the runnable semantics -- and the
oracle -- are sa1 only. sa2/sa3/fc are kept as parameter holders so reference
checkpoints load key for key. `chained=True` runs the intended three-layer
stack instead.
"""
import torch.nn as nn

from .pointnet2_utils import PointNetSetAbstraction


class feat_extraction_layer(nn.Module):
    def __init__(self, use_normal=True, npoint=10000, radius=0.1, nsample=256, chained=False):
        super().__init__()
        in_channel = 6 if use_normal else 3
        self.use_normal = use_normal
        self.chained = chained
        self.sa1 = PointNetSetAbstraction(npoint=npoint, radius=radius, nsample=nsample, in_channel=in_channel,
                                          mlp=[16, 16, 32], group_all=False)
        self.sa2 = PointNetSetAbstraction(npoint=npoint, radius=0.2, nsample=128, in_channel=32, mlp=[32, 64],
                                          group_all=False)
        self.sa3 = PointNetSetAbstraction(npoint=npoint, radius=0.4, nsample=64, in_channel=64, mlp=[64, 64],
                                          group_all=False)
        self.fc = nn.Linear(64, 32)

    def forward(self, pts, start=None):
        """pts [B,C_in,N] -> (xyz [B,S,3], feats [B,S,32]) in FPS order."""
        if self.chained:
            raise NotImplementedError("the repaired three-layer stack is a later row (SURVEY 8f)")
        if self.use_normal:
            xyz, normal = pts[:, :3, :], pts[:, 3:, :]
        else:
            xyz, normal = pts, None
        oxyz, opts = self.sa1(xyz, normal, start=start)
        return oxyz.permute(0, 2, 1), opts.permute(0, 2, 1)
