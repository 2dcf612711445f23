"""Point-set primitives behind the reference's names and signatures
(pointnet2_utils.py:19-202 of the reference), computed by sm_100a kernels.

Only what the DeepVCP path uses is provided: `square_distance`, `index_points`,
`farthest_point_sample`, `query_ball_point`, `sample_and_group`,
`PointNetSetAbstraction` (group_all=False). Inputs must live on a CUDA device.
"""
import torch
import torch.nn as nn

from . import functional as F_
from ._lib import cloud_cm, cloud_pm, require_cuda


def _float_cloud(xyz):
    if xyz.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("float32 or float64 cloud expected (got %s)" % xyz.dtype)
    return xyz


def square_distance(src, dst):
    """[B,N,3], [B,M,3] -> [B,N,M], expanded form (reference :19-40); float32 clouds give float32, a
    float64 cloud on either side gives float64 like torch's promotion."""
    return F_.square_distance(_float_cloud(src), _float_cloud(dst))


def index_points(points, idx):
    """points [B,N,C], idx [B,S] or [B,S,K] -> [B,S,(K),C] (reference :43-60), dtype of `points`."""
    if points.dtype == torch.float64:   # a plain gather: torch's own kernel (plumbing, no arithmetic)
        B = points.shape[0]
        flat = idx.reshape(B, -1, 1).expand(-1, -1, points.shape[-1])
        return torch.gather(points, 1, flat).reshape(*idx.shape, points.shape[-1])
    return F_.index_points(points, idx)


def farthest_point_sample(xyz, npoint, start=None):
    """xyz [B,N,3] (float32 or float64) -> centroids [B,npoint] int64 (reference :63-84).

    `start` ([B] indices) replaces the reference's internal random draw; when
    omitted it is drawn exactly as the reference does (CPU default generator).
    """
    require_cuda(xyz)
    B, N, _ = xyz.shape
    if xyz.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("farthest_point_sample: float32 or float64 cloud expected")
    if start is None:
        start = F_.draw_fps_start(B, N)
    out, _ = F_.fps(cloud_pm(xyz), xyz.device, xyz.dtype, B, N, npoint, start)
    return out


def query_ball_point(radius, nsample, xyz, new_xyz):
    """[B,N,3], [B,S,3] -> group_idx [B,S,nsample] int64 (reference :87-107)."""
    B, N, _ = xyz.shape
    if N < nsample:
        raise IndexError("query_ball_point needs N >= nsample (reference :106)")
    return F_.ball_query(radius, nsample, _float_cloud(xyz), _float_cloud(new_xyz))


def sample_and_group(npoint, radius, nsample, xyz, points, returnidx=False, start=None):
    """Reference :110-138. xyz [B,N,3], points [B,N,D] or None."""
    B, N, C = xyz.shape
    fps_idx = farthest_point_sample(xyz, npoint, start)
    new_xyz = index_points(xyz, fps_idx)
    idx = query_ball_point(radius, nsample, xyz, new_xyz)
    if int(idx.max()) >= N:   # empty ball: the reference raises in index_points
        raise IndexError("index %d is out of bounds for dimension 1 with size %d" % (int(idx.max()), N))
    grouped_xyz_norm = index_points(xyz, idx) - new_xyz.view(B, npoint, 1, C)
    if points is not None:
        new_points = torch.cat([grouped_xyz_norm, index_points(points, idx)], dim=-1)
    else:
        new_points = grouped_xyz_norm
    if returnidx:
        return new_xyz, new_points, idx
    return new_xyz, new_points


class PointNetSetAbstraction(nn.Module):
    """Reference :161-202 with identical parameters / state_dict keys. forward()
    runs FPS, then ONE fused kernel (ball query + grouping + shared MLP with
    eval-mode BatchNorm + max over the ball); the [B,3+D,nsample,npoint] tensor of
    the reference is never built."""

    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv2d(last, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out_channel))
            last = out_channel
        self.group_all = group_all
        self._folded = None
        self._folded_key = None

    def folded(self):
        key = tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))
        if self._folded is None or key != self._folded_key:
            self._folded = F_.FoldedMlp(list(self.mlp_convs), list(self.mlp_bns))
            self._folded_key = key
        return self._folded

    def forward(self, xyz, points, start=None, return_fps=False):
        """xyz [B,3,N], points [B,D,N] or None -> new_xyz [B,3,S], new_points [B,D',S]."""
        if self.group_all:
            raise NotImplementedError("group_all=True is not on the DeepVCP path")
        if self.training:
            # batch statistics + autograd graph: sampling and ball query on the kernels, the shared MLP with
            # torch operators (training.py; reference :176-202 in train mode)
            from . import training
            new_xyz, feats, fps = training.set_abstraction(self, xyz, points, start)
            out = (new_xyz.permute(0, 2, 1), feats.permute(0, 2, 1))
            return out + (fps.int(),) if return_fps else out
        require_cuda(xyz, points)
        _float_cloud(xyz)
        B, _, N = xyz.shape
        S = self.npoint
        dev = xyz.device
        if start is None:
            start = F_.draw_fps_start(B, N)
        if xyz.dtype == torch.float64 or (points is not None and points.dtype == torch.float64):
            # torch promotes the grouped tensor to double and casts right before the MLP (reference :128-132,198)
            xyz64 = xyz.double()
            D = 0 if points is None else points.shape[1]
            pts64 = points.double() if points is not None else None
            fps64, fps32 = F_.fps(cloud_cm(xyz64), dev, xyz64.dtype, B, N, S, start, want64=return_fps, want32=True)
            new_xyz, feats = F_.sa_layer_f64(cloud_cm(xyz64), cloud_cm(pts64) if D else None, D, fps32, B, N, S,
                                             self.radius, self.nsample, self.folded(), dev)
            out = (new_xyz.to(xyz.dtype).permute(0, 2, 1), feats.permute(0, 2, 1))
            return out + (fps32,) if return_fps else out
        index = F_.SpatialIndex(B, N, dev) if F_.SpatialIndex.indexable(N) else None
        _, fps32 = F_.fps(cloud_cm(xyz), dev, xyz.dtype, B, N, S, start, want64=return_fps, want32=True, index=index)
        D = 0 if points is None else points.shape[1]
        if points is not None and points.dtype != torch.float32:
            points = points.float()
        new_xyz, feats = F_.sa_layer(cloud_cm(xyz), cloud_cm(points) if D else None, D, fps32, B, N, S,
                                     self.radius, self.nsample, self.folded(), dev, index=index)
        out = (new_xyz.permute(0, 2, 1), feats.permute(0, 2, 1))
        return out + (fps32,) if return_fps else out
