"""Seeded synthetic point-cloud pairs in the two shapes BASELINE.json names.

There is no dataset in the build or bench environment, so both the parity tests
and ``bench.py`` draw their inputs here. The generators mirror what the
reference's loaders hand to the model (shape, layout, dtype, value ranges):

* ModelNet40-shaped  -- ``ModelNet40Dataset.py:38-41,62-92`` of the reference:
  unit-ball cloud with unit normals, ``[6, N]`` channel-major, random
  ``R = RotX RotY RotZ`` and ``t ~ U[-1,1]^3``, ``tgt = R src + t``.
* KITTI-shaped       -- ``KITTIDataset.py:39-46,67-84``: xyz only ``[3, N]``;
  a ring-like scan snapped to a 0.1 m lattice and de-duplicated so that exact
  distance ties occur and the tie-breaking rules are exercised.

Everything is drawn from a CPU ``torch.Generator`` so a pair is a pure function
of ``pair_id``.
"""
import math

import torch


def _rot_xyz(ax: float, ay: float, az: float) -> torch.Tensor:
    """RotX(ax) @ RotY(ay) @ RotZ(az) in float64 (reference utils.py:8-26)."""
    cx, sx = math.cos(ax), math.sin(ax)
    cy, sy = math.cos(ay), math.sin(ay)
    cz, sz = math.cos(az), math.sin(az)
    rx = torch.tensor([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=torch.float64)
    ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=torch.float64)
    rz = torch.tensor([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], dtype=torch.float64)
    return rx @ ry @ rz


def _random_pose(g: torch.Generator, max_angle: float = 2 * math.pi):
    ang = torch.rand(3, generator=g, dtype=torch.float64) * max_angle
    R = _rot_xyz(float(ang[0]), float(ang[1]), float(ang[2]))
    t = torch.rand(3, generator=g, dtype=torch.float64) * 2 - 1
    return R, t


def _cast(src64, tgt64, dtype):
    """dtype "f32": what every benchmark uses; "f64": as ModelNet40Dataset.py:38,92 hands the clouds over
    (float64 arrays); "mixed": as KITTIDataset.py:84,97 does (float32 scan, target = R @ src + t in float64)."""
    if dtype == "f32":
        return src64.float(), tgt64.float()
    if dtype == "f64":
        return src64, tgt64
    if dtype == "mixed":
        return src64.float(), tgt64
    raise ValueError("dtype must be f32, f64 or mixed")


def modelnet_pair(pair_id: int, n_points: int = 1024, dtype: str = "f32"):
    """One ModelNet40-shaped pair.

    Returns ``src[6,N], tgt[6,N] (float32 by default), R[3,3] f64, t[3] f64``.
    """
    g = torch.Generator().manual_seed(1234 + int(pair_id))
    d = torch.randn(n_points, 3, generator=g, dtype=torch.float64)
    d = d / d.norm(dim=1, keepdim=True)
    rad = (torch.rand(n_points, 1, generator=g, dtype=torch.float64) * 0.8 + 0.2) ** (1.0 / 3.0)
    xyz = d * rad
    R, t = _random_pose(g)
    tgt_xyz = xyz @ R.T + t
    tgt_nrm = d @ R.T
    src, tgt = _cast(torch.cat([xyz, d], dim=1).T.contiguous(), torch.cat([tgt_xyz, tgt_nrm], dim=1).T.contiguous(), dtype)
    return src, tgt, R, t


def _kitti_scan(g: torch.Generator, n_points: int) -> torch.Tensor:
    """n_points unique lattice points (0.1 m) of a synthetic lidar-like scan, float64."""
    keys = set()
    rows = []
    need = n_points
    while need > 0:
        m = max(2 * need, 1024)
        rho = (torch.randn(m, generator=g, dtype=torch.float64) * 25.0).abs().clamp(max=80.0)
        az = torch.rand(m, generator=g, dtype=torch.float64) * (2 * math.pi)
        u = torch.rand(m, generator=g, dtype=torch.float64)
        v = torch.rand(m, generator=g, dtype=torch.float64)
        z = torch.where(u < 0.7, v * 3.0 - 2.0, v * 5.0 + 1.0)
        p = torch.stack([rho * torch.cos(az), rho * torch.sin(az), z], dim=1)
        q = torch.round(p * 10.0).to(torch.int64)
        for k in range(m):
            key = (int(q[k, 0]), int(q[k, 1]), int(q[k, 2]))
            if key in keys:
                continue
            keys.add(key)
            rows.append(key)
            need -= 1
            if need == 0:
                break
    return torch.tensor(rows, dtype=torch.float64) / 10.0


def kitti_pair(pair_id: int, n_points: int = 16384, max_angle: float = 2 * math.pi, dtype: str = "f32"):
    """One KITTI-shaped pair: ``src[3,N], tgt[3,N] (float32 by default), R f64, t f64``."""
    g = torch.Generator().manual_seed(4321 + int(pair_id))
    xyz = _kitti_scan(g, n_points)
    R, t = _random_pose(g, max_angle)
    if dtype == "mixed":   # the loader transforms the float32 scan in float64 (KITTIDataset.py:84)
        xyz = xyz.float().double()
    tgt_xyz = xyz @ R.T + t
    src, tgt = _cast(xyz.T.contiguous(), tgt_xyz.T.contiguous(), dtype)
    return src, tgt, R, t


def make_batch(kind: str, pair_ids, n_points: int, dtype: str = "f32"):
    """Stack pairs: ``src[B,C,N], tgt[B,C,N] (float32 unless dtype says otherwise), R[B,3,3] f64, t[B,3] f64``."""
    fn = {"modelnet": modelnet_pair, "kitti": kitti_pair}[kind]
    items = [fn(i, n_points, dtype=dtype) for i in pair_ids]
    src = torch.stack([it[0] for it in items])
    tgt = torch.stack([it[1] for it in items])
    R = torch.stack([it[2] for it in items])
    t = torch.stack([it[3] for it in items])
    return src, tgt, R, t


def grid_radius(G: int, s: float = 0.4) -> float:
    """Search radius r for which both ``voxelize`` and ``cpg`` see a G^3 grid.

    ``cpg`` takes ``int(2r/s + 1)`` (cpg.py:29) while ``voxelize`` takes the length of
    an ``arange`` (voxelize.py:62-64); r = (G-1)*s/2 satisfies both for odd G in
    5..21 with s = 0.4, and r = 1.0 gives the reference's own 6^3.
    """
    return (G - 1) * s / 2.0
