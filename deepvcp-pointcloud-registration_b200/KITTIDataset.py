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
