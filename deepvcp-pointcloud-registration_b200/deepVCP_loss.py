"""Pose solve and loss -- deepVCP_loss.py:13-121 of the reference. The kernels compute values; a prediction
that carries a gradient goes through the autograd form of training.py."""
import torch

from . import functional as F_
from ._lib import QUIRKS_REFERENCE


def get_rigid_transform(x, y, quirks=QUIRKS_REFERENCE, weights=None):
    """x, y [B,3,N] -> R [B,3,3], t [B,3,1] (reference :13-44): Kabsch with
    R = V U^T and, in reference mode, no reflection correction (quirk Q10; pass
    quirks without QUIRK_NO_REFLECTION_FIX for det R = +1). weights [B,N]: the
    weighted solve. Computed in float64 and returned in the input dtype."""
    R, t = F_.kabsch(x, y, quirks=quirks, weights=weights)
    return R.to(x.dtype), t.to(x.dtype)


def svd_optimization(x, y_pred, R_true, t_true, quirks=QUIRKS_REFERENCE):
    """Reference :57-90: solve, drop the 20 % of points with the largest 1-NN
    distance to the ground-truth transform of x, solve again.
    Returns R2, t2, x1 [B,3,n'], y_pred2 [B,3,n'] like the reference: x1 = the inliers of x in the order of
    its topk(largest=False, sorted=True) (:77,82), y_pred2 = R2 x1 + t2 (:88)."""
    R2, t2, _, _, inl = F_.kabsch_refine(x, y_pred, R_true, t_true, quirks=quirks, want_inliers=True)
    x1 = torch.gather(x.double(), 2, inl.unsqueeze(1).expand(-1, 3, -1))
    y_pred2 = torch.matmul(R2, x1) + t2
    return R2, t2, x1, y_pred2


def deepVCP_loss(x, y_pred, R_true, t_true, alpha):
    """Reference :105-121 (a y_pred with requires_grad takes the differentiable path of training.py). x, y_pred [B,N,3]; returns
    (loss, R [B,3,3], t [B,3,1]): alpha * L1(y_true_inliers, y_pred2) + (1 - alpha) * |mean(y_pred2 - y_true_inliers)|."""
    if y_pred.requires_grad:
        from . import training
        return training.loss(x, y_pred, R_true, t_true, alpha)
    xx = x.permute(0, 2, 1).double()
    yy = y_pred.permute(0, 2, 1).double()
    R, t, x_inl, y_opt = svd_optimization(xx, yy, R_true, t_true)
    y_true_inl = torch.matmul(R_true.double(), x_inl) + t_true.double().reshape(x.shape[0], 3, -1)
    loss1 = torch.mean(torch.abs(y_true_inl - y_opt))
    loss2 = torch.abs(torch.mean(y_opt - y_true_inl))
    return alpha * loss1 + (1 - alpha) * loss2, R, t


def pose_from_forward(src_keypts, tgt_vcp, R_true, t_true, quirks=QUIRKS_REFERENCE):
    """train.py:110 -> deepVCP_loss.py:105-107,121: permute to [B,3,N], double,
    two-stage solve -- done inside one kernel that reads the forward's float32
    [B,N,3] outputs in place. Returns R [B,3,3], t [B,3,1] float64."""
    if src_keypts.dtype == torch.float32 and tgt_vcp.dtype == torch.float32:
        return F_.pose_from_forward(src_keypts, tgt_vcp, R_true, t_true, quirks=quirks)
    x = src_keypts.permute(0, 2, 1).double()
    y = tgt_vcp.permute(0, 2, 1).double()
    R2, t2, _, _ = F_.kabsch_refine(x, y, R_true, t_true, quirks=quirks)
    return R2, t2
