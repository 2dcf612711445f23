"""Get_Cat_Feat_Tgt -- get_cat_feat_tgt.py:14-98 of the reference.

Standalone module: KNN kernel + gathers; returns the float64
[B,K_topk,C,k_nn,3+num_feat] tensor like the reference. The DeepVCP forward does
NOT call this: it uses the fused gather+embedding kernel and never builds the
tensor (763 MB per KITTI-shaped pair)."""
import torch
import torch.nn as nn

from . import functional as F_
from ._lib import cloud_pm, require_cuda


class Get_Cat_Feat_Tgt(nn.Module):
    def __init__(self, k_nn=32, per_feature_weight=True):
        super().__init__()
        self.k_nn = k_nn
        self.per_feature_weight = per_feature_weight   # quirk Q7

    def forward(self, candidate_pts, src_keypts, tgt_pts_xyz, tgt_deep_feat_pts):
        require_cuda(candidate_pts, tgt_pts_xyz, tgt_deep_feat_pts)
        B, M, C, _ = candidate_pts.shape
        N = tgt_pts_xyz.shape[1]
        K = self.k_nn
        Fd = tgt_deep_feat_pts.shape[2]
        cand = candidate_pts.float().contiguous()
        ref = tgt_pts_xyz.float()
        dist, idx, _ = F_.knn(cloud_pm(ref), ref.device, B, N, cand.view(B, M * C, 3), K)
        w = dist / dist.sum(dim=2, keepdim=True, dtype=torch.float64)
        feat = F_.index_points(tgt_deep_feat_pts.float(), idx).view(B, M, C, K, Fd)
        xyz = F_.index_points(ref.contiguous(), idx).view(B, M, C, K, 3)
        local = xyz - cand.unsqueeze(3)
        if self.per_feature_weight:
            assert K == Fd, "k_nn must equal the feature width (reference :92)"
            wmap = w.view(B, M, C, 1, K)
        else:
            wmap = w.view(B, M, C, K, 1)
        return torch.cat((local, feat * wmap), dim=4)
