"""ModelNet40-shaped data ingest -- ModelNet40Dataset.py:12-92 of the reference.

The reference parses `<root>/<category>/<name>.txt` (comma-separated x, y, z, nx, ny, nz) with np.loadtxt into
float64 arrays and, per item, draws three angles and a translation, forms R = RotX RotY RotZ (utils.py:8-26) and
builds `target = (R @ xyz + t, R @ normals)` in numpy. Text parsing stays on the host; the clouds then go to the
device once and ONE kernel writes the source and the transformed target in the model's [B, 6, N] layout, in
float64 (what the reference's loader yields, :38,92) or float32 (what the fast kernels take). The random draws
are made exactly like the reference makes them (np.random.uniform x 3, then torch.rand(3, 1)), so a seeded run
produces the same pairs."""
import math
import os

import numpy as np
import torch

from . import functional as F_
from ._lib import check, lib, ptr, stream_ptr


def read_txt(path: str) -> np.ndarray:
    """ModelNet40Dataset.py:38: one cloud as [M, 6] float64 (x, y, z, nx, ny, nz)."""
    return np.loadtxt(path, delimiter=",", dtype=np.float64)


def rotation(theta_x, theta_y, theta_z) -> np.ndarray:
    """R = RotX(theta_x) @ RotY(theta_y) @ RotZ(theta_z) (utils.py:8-26, ModelNet40Dataset.py:72-75) as [3,3] float64."""
    cx, sx, cy, sy, cz, sz = (math.cos(theta_x), math.sin(theta_x), math.cos(theta_y), math.sin(theta_y),
                              math.cos(theta_z), math.sin(theta_z))
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=np.float64)
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=np.float64)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], dtype=np.float64)
    return Rx @ Ry @ Rz


def draw_transform():
    """The draws of one __getitem__ in the reference's order (:62-69): three np.random.uniform angles, then
    torch.rand(3, 1) for the translation in [-1, 1). Returns R [3,3] float64 (numpy), t [3,1] float32 (torch)."""
    tx, ty, tz = (np.random.uniform(0, np.pi * 2) for _ in range(3))
    t = (1.0 - -1.0) * torch.rand(3, 1) + -1.0
    return rotation(tx, ty, tz), t


def ingest(clouds, R=None, t=None, N=None, device="cuda", dtype=torch.float64):
    """clouds: B arrays [M, 6] float64 (equal M); R [B,3,3], t [B,3] (None: no target); N rows kept (default M).
    Returns src [B,6,N], tgt [B,6,N] or None on the device in `dtype` (float64: the reference's; float32)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("ingest runs on a CUDA device; there is no CPU fallback")
    raw = torch.stack([torch.as_tensor(np.asarray(c), dtype=torch.float64).reshape(-1, 6) for c in clouds])
    B, M, _ = raw.shape
    N = M if N is None else int(N)
    if dtype not in (torch.float64, torch.float32):
        raise RuntimeError("float64 or float32")
    raw = raw.to(dev, non_blocking=True).contiguous()
    src = torch.empty(B, 6, N, dtype=dtype, device=dev)
    tgt = torch.empty(B, 6, N, dtype=dtype, device=dev) if R is not None else None
    Rd = torch.as_tensor(np.asarray(R), dtype=torch.float64).reshape(B, 9).to(dev).contiguous() if R is not None else None
    td = torch.as_tensor(np.asarray(t), dtype=torch.float64).reshape(B, 3).to(dev).contiguous() if R is not None else None
    check(lib().dvcp_ingest_modelnet(ptr(raw), ptr(Rd), ptr(td), B, M, N, int(dtype == torch.float64), ptr(src), ptr(tgt),
                                     stream_ptr(dev)), "dvcp_ingest_modelnet")
    F_._count(1)
    return src, tgt


class ModelNet40Dataset(torch.utils.data.Dataset):
    """Same constructor arguments and item layout as the reference's class (:12-92): item = (src [6,N],
    target [6,N], R [3,3] float64, t [3,1]); the clouds are device tensors (float64 by default)."""

    def __init__(self, root, augment=True, rotate=True, full_dataset=True, split="train", device="cuda",
                 dtype=torch.float64):
        self.root, self.split, self.augment = root, split, augment
        self.device, self.dtype = device, dtype
        self.cat = [line.rstrip() for line in open(os.path.join(root, "modelnet10_shape_names.txt"))]
        listing = "modelnet10_%s.txt" % split if full_dataset else "modelnet10_small_%s.txt" % split
        names = np.loadtxt(os.path.join(root, listing), dtype=str).reshape(-1)
        self.clouds, self.labels = [], []
        for name in names:
            category, _ = name.split("_0")
            self.clouds.append(read_txt(os.path.join(root, category, name) + ".txt"))
            self.labels.append(name)

    def __len__(self):
        return len(self.clouds)

    def __getitem__(self, index):
        if not self.augment:   # the reference leaves target undefined here (:59,84): identity transform instead
            R, t = np.eye(3), torch.zeros(3, 1)
        else:
            R, t = draw_transform()
        src, tgt = ingest([self.clouds[index]], R[None], t.reshape(1, 3).double(), device=self.device, dtype=self.dtype)
        return src[0], tgt[0], torch.from_numpy(R), t
