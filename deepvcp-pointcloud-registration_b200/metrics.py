"""Registration error metrics as the reference's training / test loops print them (train.py:113-120,158-165):
rotation error = L2 distance of the Euler-xyz angle triples (degrees) of the predicted and the true rotation,
translation error = L2 distance of the translations. Host-side arithmetic on [B,3,3] / [B,3,1] results (the
reference itself goes through scipy on the CPU here); nothing of the hot path."""
import numpy as np
import torch


def euler_xyz_degrees(R: torch.Tensor) -> torch.Tensor:
    """[B,3,3] (or [3,3]) -> [B,3] extrinsic x-y-z Euler angles in degrees = scipy's Rotation.as_euler('xyz',
    degrees=True) convention (train.py:114-117)."""
    from scipy.spatial.transform import Rotation
    M = R.detach().double().cpu().reshape(-1, 3, 3).numpy()
    out = np.stack([Rotation.from_matrix(m).as_euler("xyz", degrees=True) for m in M])
    return torch.from_numpy(out)


def registration_errors(R_pred, t_pred, R_gt, t_gt):
    """-> (rotation error [B], translation error [B]) exactly as train.py:113-120 computes them for B = 1."""
    B = R_pred.reshape(-1, 3, 3).shape[0]
    rot = (euler_xyz_degrees(R_pred) - euler_xyz_degrees(R_gt)).norm(dim=1)
    d = t_pred.detach().double().cpu().reshape(B, 3) - t_gt.detach().double().cpu().reshape(B, 3)
    return rot, (d + 1e-6).norm(dim=1)          # nn.PairwiseDistance adds eps = 1e-6 to the difference
