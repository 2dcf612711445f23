"""feat_extraction_layer -- deep_feat_extraction.py:5-32 of the reference.

The reference's forward crashes after sa1 at HEAD (it feeds normals into layers
declared for 32/64 channels); the runnable semantics -- and the
oracle -- are sa1 only. sa2/sa3/fc are kept as parameter holders so reference
checkpoints load key for key.

`chained=True` builds and runs the three-layer stack the file intends (SURVEY 8f
rank 1): sa1 -> sa2 -> sa3 with each layer's xyz AND features handed to the next,
then `fc`. A set-abstraction layer consumes [xyz - centre, features], so sa2 / sa3
need in_channel = 3 + 32 / 3 + 64 (the reference declares 32 / 64, one reason it
cannot run): a chained model has those two weight shapes and does not load a
reference checkpoint key for key. Every layer samples with its own FPS start.
"""
import torch
import torch.nn as nn

from . import functional as F_
from .pointnet2_utils import PointNetSetAbstraction


class feat_extraction_layer(nn.Module):
    def __init__(self, use_normal=True, npoint=10000, radius=0.1, nsample=256, chained=False):
        super().__init__()
        in_channel = 6 if use_normal else 3
        self.use_normal = use_normal
        self.chained = chained
        extra = 3 if chained else 0
        self.sa1 = PointNetSetAbstraction(npoint=npoint, radius=radius, nsample=nsample, in_channel=in_channel,
                                          mlp=[16, 16, 32], group_all=False)
        self.sa2 = PointNetSetAbstraction(npoint=npoint, radius=0.2, nsample=128, in_channel=32 + extra, mlp=[32, 64],
                                          group_all=False)
        self.sa3 = PointNetSetAbstraction(npoint=npoint, radius=0.4, nsample=64, in_channel=64 + extra, mlp=[64, 64],
                                          group_all=False)
        self.fc = nn.Linear(64, 32)

    def forward(self, pts, start=None, return_fps=False):
        """pts [B,C_in,N] -> (xyz [B,S,3], feats [B,S,32]) in FPS order. start: [B] FPS start of sa1, or
        [3,B] for the three layers of the chained stack (missing ones are drawn like the reference draws)."""
        if self.use_normal:
            xyz, normal = pts[:, :3, :], pts[:, 3:, :]
        else:
            xyz, normal = pts, None
        if not self.chained:
            out = self.sa1(xyz, normal, start=start, return_fps=return_fps)
            res = (out[0].permute(0, 2, 1), out[1].permute(0, 2, 1))
            return res + (out[2],) if return_fps else res
        st = [None, None, None]
        if start is not None:
            start = torch.as_tensor(start)
            if start.dim() == 2:
                st = [start[0], start[1], start[2]]
            else:
                st[0] = start
        x1, f1 = self.sa1(xyz, normal, start=st[0])
        x2, f2 = self.sa2(x1, f1, start=st[1])
        x3, f3, fps3 = self.sa3(x2, f2, start=st[2], return_fps=True)
        feats = F_.linear_rows(f3.permute(0, 2, 1), self.fc.weight, self.fc.bias)     # [B,S,32]
        res = (x3.permute(0, 2, 1), feats)
        return res + (fps3,) if return_fps else res
