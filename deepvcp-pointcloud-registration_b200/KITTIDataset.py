"""KITTI-shaped data ingest on the GPU -- KITTIDataset.py:11-16,39-46,67-84 of the reference
(the step immediately before the hot path).

The reference reads a velodyne `.bin` file into an [M, 4] float32 array, keeps N random rows
(`downsample`), splits xyz / reflectance and builds the target cloud `R @ src + t` in numpy. Here the
raw scans go to the device once and ONE kernel gathers the chosen rows into the model's [B, 3, N]
layout and writes the transformed target beside it. The random choices stay on the host, drawn exactly
like the reference draws them (`np.random.choice(num_src, N, replace=False)`), so a seeded run picks
the same points."""
import numpy as np
import torch

from . import functional as F_
from ._lib import check, lib, ptr, stream_ptr


def downsample_indices(num_src: int, N: int, rng=np.random):
    """KITTIDataset.py:11-16: the row indices `downsample` keeps (all rows when num_src <= N)."""
    if num_src > N:
        return rng.choice(num_src, N, replace=False)
    return np.arange(num_src)


def read_velodyne(path: str) -> np.ndarray:
    """KITTIDataset.py:39: one scan as [M, 4] float32 (x, y, z, reflectance)."""
    return np.fromfile(path, dtype=np.float32, count=-1).reshape([-1, 4])


def ingest(scans, idx, R=None, t=None, device="cuda", want_reflectance=False):
    """scans: list of B [M_b, 4] float32 arrays / tensors (host or device); idx [B, N] int64 rows to
    keep (None: the first N = min M_b rows); R [B,3,3], t [B,3] float64 (None: no target).
    Returns src [B,3,N] f32, tgt [B,3,N] f32 or None, reflectance [B,1,N] or None -- device tensors."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("ingest runs on a CUDA device; there is no CPU fallback")
    B = len(scans)
    ts = [torch.as_tensor(s, dtype=torch.float32).reshape(-1, 4) for s in scans]
    rows = [int(s.shape[0]) for s in ts]
    offs = torch.tensor([0] + list(np.cumsum(rows)), dtype=torch.int64)
    raw = torch.cat([s.to(dev, non_blocking=True) for s in ts], 0).contiguous()
    if idx is None:
        N = min(rows)
        idx_d = None
    else:
        ia = np.asarray(idx).reshape(B, -1)
        for b in range(B):
            if ia[b].min() < 0 or ia[b].max() >= rows[b]:
                raise IndexError("ingest: row index outside scan %d (%d rows)" % (b, rows[b]))
        idx_d = torch.as_tensor(ia, dtype=torch.int64).to(dev, non_blocking=True).contiguous()
        N = idx_d.shape[1]
    src = torch.empty(B, 3, N, dtype=torch.float32, device=dev)
    tgt = torch.empty(B, 3, N, dtype=torch.float32, device=dev) if R is not None else None
    refl = torch.empty(B, 1, N, dtype=torch.float32, device=dev) if want_reflectance else None
    Rd = torch.as_tensor(R, dtype=torch.float64).reshape(B, 9).to(dev).contiguous() if R is not None else None
    td = torch.as_tensor(t, dtype=torch.float64).reshape(B, 3).to(dev).contiguous() if R is not None else None
    check(lib().dvcp_ingest_kitti(ptr(raw), ptr(offs.to(dev)), ptr(idx_d), ptr(Rd), ptr(td), B, N, ptr(src), ptr(tgt),
                                  ptr(refl), stream_ptr(dev)), "dvcp_ingest_kitti")
    F_._count(1)
    return src, tgt, refl


def voxel_grid_filter(points, cell, origin=(0.0, 0.0, 0.0), mode="centroid", capacity=None, want_counts=False):
    """Voxel-grid filter on the GPU (the alternative to the random down-sample of
    KITTIDataset.py:11-16 that the paper's KITTI pipeline uses). points [M, 3 or 4] float32 (host or device):
    one output row per occupied cell of edge `cell` -- the centroid of the cell's points (mode "centroid",
    float64 sums in point order) or its first point (mode "first"); cells come out in (ix, iy, iz) order.
    Returns a device tensor [n_cells, C] (and the points per cell, int32, if want_counts)."""
    pts = torch.as_tensor(points, dtype=torch.float32)
    if pts.dim() != 2 or pts.shape[1] not in (3, 4):
        raise RuntimeError("voxel_grid_filter: [M,3] or [M,4] float32 points expected")
    if not pts.is_cuda:
        pts = pts.to("cuda") if torch.cuda.is_available() else pts
    if not pts.is_cuda:
        raise RuntimeError("voxel_grid_filter runs on a CUDA device; there is no CPU fallback")
    pts = pts.contiguous()
    dev = pts.device
    M, C = pts.shape
    cap = int(capacity) if capacity else M
    ws = torch.empty(int(lib().dvcp_voxel_filter_workspace_bytes(M)), dtype=torch.uint8, device=dev)
    out = torch.empty(cap, C, dtype=torch.float32, device=dev)
    cnt = torch.empty(cap, dtype=torch.int32, device=dev) if want_counts else None
    n_out = torch.zeros(1, dtype=torch.int64, device=dev)
    check(lib().dvcp_voxel_grid_filter(ptr(pts), C, C, M, float(origin[0]), float(origin[1]), float(origin[2]), float(cell),
                                       {"centroid": 0, "first": 1}[mode], ptr(ws), cap, ptr(out), ptr(cnt), ptr(n_out),
                                       stream_ptr(dev)), "dvcp_voxel_grid_filter")
    F_._count(5)
    n = min(int(n_out.item()), cap)
    return (out[:n], cnt[:n]) if want_counts else out[:n]
