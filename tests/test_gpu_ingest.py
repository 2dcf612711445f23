"""Data ingest on the GPU: the ModelNet40-shaped loader (ModelNet40Dataset.py:34-46,62-92) and
the voxel-grid filter, against the reference loader's numpy arithmetic and the oracle's restatement."""
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import PKG
from oracle import native

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def dv():
    return importlib.import_module(PKG)


def write_modelnet_tree(root, n_clouds=3, rows=300, seed=0):
    rng = np.random.RandomState(seed)
    os.makedirs(os.path.join(root, "chair"), exist_ok=True)
    names = ["chair_%04d" % (i + 1) for i in range(n_clouds)]
    open(os.path.join(root, "modelnet10_shape_names.txt"), "w").write("chair\n")
    open(os.path.join(root, "modelnet10_train.txt"), "w").write("\n".join(names) + "\n")
    clouds = []
    for n in names:
        xyz = rng.uniform(-1, 1, (rows, 3))
        nrm = rng.normal(size=(rows, 3))
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        data = np.concatenate([xyz, nrm], 1)
        with open(os.path.join(root, "chair", n + ".txt"), "w") as f:
            for r in data:
                f.write(",".join("%.6f" % v for v in r) + "\n")
        clouds.append(np.loadtxt(os.path.join(root, "chair", n + ".txt"), delimiter=",", dtype=np.float64))
    return names, clouds


def test_modelnet_dataset_items_follow_the_reference_loader(dv, tmp_path):
    root = str(tmp_path)
    names, clouds = write_modelnet_tree(root)
    ds = dv.ModelNet40Dataset.ModelNet40Dataset(root, augment=True, split="train")
    assert len(ds) == 3 and ds.labels == names
    for index in range(3):
        np.random.seed(100 + index)
        torch.manual_seed(200 + index)
        src, tgt, R, t = ds[index]
        # the reference's __getitem__ (ModelNet40Dataset.py:54-92), restated with numpy on the same draws
        np.random.seed(100 + index)
        torch.manual_seed(200 + index)
        th = [np.random.uniform(0, np.pi * 2) for _ in range(3)]
        t_ref = (1.0 - -1.0) * torch.rand(3, 1) + -1.0
        c, s = np.cos, np.sin
        Rx = np.array([[1, 0, 0], [0, c(th[0]), -s(th[0])], [0, s(th[0]), c(th[0])]])
        Ry = np.array([[c(th[1]), 0, s(th[1])], [0, 1, 0], [-s(th[1]), 0, c(th[1])]])
        Rz = np.array([[c(th[2]), -s(th[2]), 0], [s(th[2]), c(th[2]), 0], [0, 0, 1]])
        R_ref = Rx @ Ry @ Rz
        P, Nn = clouds[index][:, :3].T, clouds[index][:, 3:].T
        tgt_ref = torch.cat((torch.from_numpy(R_ref @ P) + t_ref, torch.from_numpy(R_ref @ Nn)), 0)
        src_ref = torch.cat((torch.from_numpy(P), torch.from_numpy(Nn)), 0)
        assert src.dtype == torch.float64 and tgt.dtype == torch.float64 and src.shape == (6, 300)   # :38,92
        assert torch.equal(t, t_ref) and np.allclose(R.numpy(), R_ref, rtol=0, atol=1e-15)
        assert torch.equal(src.cpu(), src_ref)
        assert torch.allclose(tgt.cpu(), tgt_ref, rtol=0, atol=1e-14)
    # float32 output: what the fast kernels take
    s32, t32 = dv.ModelNet40Dataset.ingest(clouds, np.stack([np.eye(3)] * 3), np.zeros((3, 3)), dtype=torch.float32)
    assert s32.dtype == torch.float32 and torch.equal(s32, t32)
    # straight into the model
    model = dv.DeepVCP(use_normal=True, npoint=300, r=0.8, s=0.4).to(DEV).eval()
    kp, vcp = model(src[None], tgt[None], R[None].to(DEV), torch.zeros(1, 3))
    assert torch.isfinite(vcp).all() and kp.dtype == torch.float64


@pytest.mark.parametrize("mode", ["centroid", "first"])
@pytest.mark.parametrize("C", [3, 4])
def test_voxel_grid_filter_vs_oracle(dv, synthetic, mode, C):
    g = torch.Generator().manual_seed(5)
    # a scan-like cloud with many points per cell near the sensor and ties on cell borders
    M = 60000
    rho = torch.randn(M, generator=g).abs() * 20
    az = torch.rand(M, generator=g) * 6.2831853
    pts = torch.stack([rho * torch.cos(az), rho * torch.sin(az), torch.rand(M, generator=g) * 6 - 2,
                       torch.rand(M, generator=g)], 1)[:, :C].contiguous()
    pts[::7, :3] = torch.round(pts[::7, :3] / 0.3) * 0.3          # points exactly on cell borders
    for cell, origin in ((0.3, (0.0, 0.0, 0.0)), (1.0, (-3.25, 0.5, 0.125))):
        out, cnt = dv.KITTIDataset.voxel_grid_filter(pts, cell, origin, mode, want_counts=True)
        ref, rcnt = native.voxel_filter(pts, cell, origin, mode)
        assert out.shape == ref.shape and torch.equal(cnt.cpu(), rcnt)
        assert torch.equal(out.cpu(), ref)                                       # bit-exact, same cell order
        assert int(cnt.sum()) == M
        # every output lies in its cell; cells are unique and ascending
        o3 = torch.tensor(origin)
        ijk = torch.floor((out.cpu()[:, :3] - o3) / cell).long() if mode == "first" else None
        if ijk is not None:
            key = (ijk[:, 0] * (1 << 42)) + (ijk[:, 1] + (1 << 20)) * (1 << 21) + (ijk[:, 2] + (1 << 20))
            assert bool((key[1:] > key[:-1]).all())
    # capacity smaller than the number of cells: the count is still reported through the tensor length cap
    few = dv.KITTIDataset.voxel_grid_filter(pts, 0.3, capacity=100)
    assert few.shape == (100, C) and torch.equal(few.cpu(), native.voxel_filter(pts, 0.3)[0][:100])


def test_voxel_filtered_scan_feeds_the_registration(dv):
    g = torch.Generator().manual_seed(9)
    scan = torch.cat([torch.randn(40000, 2, generator=g) * 15, torch.rand(40000, 1, generator=g) * 4 - 2,
                      torch.rand(40000, 1, generator=g)], 1)
    vox = dv.KITTIDataset.voxel_grid_filter(scan, 0.4)
    N = 4096
    assert vox.shape[0] >= N
    R = torch.eye(3, dtype=torch.float64)[None]
    src, tgt, _ = dv.KITTIDataset.ingest([vox], np.arange(N)[None], R, torch.tensor([[0.3, -0.2, 0.1]], dtype=torch.float64))
    model = dv.DeepVCP(use_normal=False, npoint=N, r=0.8, s=0.4).to(DEV).eval()
    kp, vcp = model(src, tgt, R.to(DEV), torch.zeros(1, 3))
    assert torch.isfinite(vcp).all()
