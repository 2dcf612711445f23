"""Get_Cat_Feat_Src -- get_cat_feat_src.py:12-55 of the reference.
Tiny ([B,64,32,35]); expressed with torch ops on the device in the input dtype.
The fused forward computes the same thing inside the key-point kernel."""
import torch
import torch.nn as nn


class Get_Cat_Feat_Src(nn.Module):
    def forward(self, src_keypts, src_keypts_grouped_pts, src_keyfeats):
        kp = src_keypts[:, :, :3].unsqueeze(2)
        grp = src_keypts_grouped_pts[:, :, :, :3]
        dist = torch.linalg.vector_norm(kp - grp + 1e-6, dim=3, keepdim=True)
        norm = dist / dist.sum(dim=2, keepdim=True)
        return torch.cat((grp - kp, src_keyfeats * norm), dim=3)
