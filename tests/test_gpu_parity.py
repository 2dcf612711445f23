"""GPU parity: every CUDA entry point (through the C ABI) against the CPU oracle
and the fixtures recorded from the reference. Index-producing stages must be
bit-exact; floating-point stages carry their tolerance in the test."""
import importlib
import math

import numpy as np
import pytest
import torch

from conftest import PKG, golden_state_dict, load_golden
from oracle import stages

pytestmark = pytest.mark.gpu
T = torch.from_numpy
DEV = "cuda"

FEAT_RTOL = 1e-3      # north_star: features within 1e-3 relative (to the tensor's max magnitude)
ROT_TOL_DEG = 1e-3    # north_star: rotation within 1e-3 degrees
TRANS_TOL = 1e-4      # north_star: translation within 1e-4 m


@pytest.fixture(scope="module")
def dv():
    return importlib.import_module(PKG)


@pytest.fixture(scope="module")
def F(dv):
    return dv.functional


@pytest.fixture(scope="module")
def prim():
    return load_golden("primitives")


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rot_angle_deg(Ra, Rb):
    Ra, Rb = Ra.double().cpu(), Rb.double().cpu()
    d = Ra @ Rb.transpose(-1, -2)
    c = ((d.diagonal(dim1=-2, dim2=-1).sum(-1) - 1) / 2).clamp(-1, 1)
    # small-angle safe: use the skew part
    s = 0.5 * torch.stack([d[..., 2, 1] - d[..., 1, 2], d[..., 0, 2] - d[..., 2, 0],
                           d[..., 1, 0] - d[..., 0, 1]], -1).norm(dim=-1)
    return torch.rad2deg(torch.atan2(s, c)).max().item()


def lattice_cloud(n, seed, extent=8.0, step=0.125):
    g = torch.Generator().manual_seed(seed)
    return torch.round((torch.rand(1, n, 3, generator=g) * 2 - 1) * extent / step) * step


# ------------------------------------------------------------------ FPS ------
@pytest.mark.parametrize("n,npoint", [(64, 64), (300, 300), (1000, 1000), (1024, 1024), (1500, 700),
                                       (2048, 2048), (5000, 5000), (10000, 10000), (16384, 16384)])
def test_fps_random_cloud_bit_exact(dv, n, npoint):
    g = torch.Generator().manual_seed(n)
    xyz = torch.rand(2, n, 3, generator=g) * 4 - 2
    start = torch.tensor([n // 3, n - 1])
    ref = stages.farthest_point_sample(xyz, npoint, start)
    out = dv.farthest_point_sample(xyz.to(DEV), npoint, start)
    assert out.dtype == torch.int64
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("n", [512, 4096])
def test_fps_lattice_ties_bit_exact(dv, n):
    xyz = lattice_cloud(n, 7)
    start = torch.tensor([5])
    ref = stages.farthest_point_sample(xyz, n, start)
    out = dv.farthest_point_sample(xyz.to(DEV), n, start)
    assert torch.equal(out.cpu(), ref)


def test_fps_kitti_shaped_full_size(dv, synthetic):
    src, _, _, _ = synthetic.make_batch("kitti", [0], 16384)
    xyz = src.permute(0, 2, 1)                           # strided view, read in place
    start = torch.tensor([1234])
    ref = stages.farthest_point_sample(xyz.contiguous(), 16384, start)
    out = dv.farthest_point_sample(xyz.to(DEV), 16384, start).cpu()
    assert torch.equal(out, ref)
    assert torch.equal(out.sort()[0], torch.arange(16384).view(1, -1))   # npoint == N: a permutation


@pytest.mark.parametrize("n,npoint,extent", [(2048, 2048, 1.0), (16384, 16384, 2.0), (4096, 1000, 0.75)])
def test_fps_dense_lattice_heavy_ties_bit_exact(dv, n, npoint, extent):
    """Dense 0.125 lattice: thousands of points share each distance value, so the batched
    kernel's candidate threshold sits inside tie classes (index order decides)."""
    xyz = lattice_cloud(n, 11, extent=extent)            # duplicates included
    start = torch.tensor([n // 2])
    ref = stages.farthest_point_sample(xyz, npoint, start)
    out = dv.farthest_point_sample(xyz.to(DEV), npoint, start)
    assert torch.equal(out.cpu(), ref)


def test_fps_duplicates_and_padding_in_indexed_kernel(dv):
    """Duplicate points leave distance-0 leftovers: the argmax then stays at the lowest index
    (pointnet2_utils.py:83), also when npoint > N."""
    g = torch.Generator().manual_seed(21)
    base = torch.rand(1, 200, 3, generator=g)
    xyz = torch.cat([base, base[:, :120], base[:, 50:130]], dim=1)      # 400 points, 200 distinct
    for npoint in (150, 400, 450):
        start = torch.tensor([333])
        ref = stages.farthest_point_sample(xyz, npoint, start)
        out = dv.farthest_point_sample(xyz.to(DEV), npoint, start)
        assert torch.equal(out.cpu(), ref)


def test_fps_one_cta_per_cloud_at_full_size(dv, F, synthetic):
    """dvcp_fps_indexed, concurrent = 2: the one-CTA-per-cloud kernel the depth >= 3 pipeline samples with. At 16384
    points it runs the WIDE rounds (two exposed keys per bucket, 64 candidates per step): random, KITTI-shaped,
    dense-lattice (heavy ties), duplicated and padded clouds, npoint < N -- against the oracle and the cluster kernel."""
    lib = importlib.import_module(PKG + "._lib")
    g = torch.Generator().manual_seed(4)
    kitti, _, _, _ = synthetic.make_batch("kitti", [3], 16384)
    base = torch.rand(1, 9000, 3, generator=g) * 10
    clouds = [
        (torch.rand(1, 16384, 3, generator=g) * 4 - 2, 16384, 5000),
        (kitti[:, :3].permute(0, 2, 1).contiguous(), 16384, 77),
        (lattice_cloud(16384, 11, extent=2.0), 16384, 8192),                       # thousands of points per distance value
        (torch.cat([base, base[:, :5000]], dim=1), 14000, 13999),                  # 14000 points (padded index), 9000 distinct
        (torch.rand(1, 16384, 3, generator=g) * 30, 700, 0),                       # npoint < N
    ]
    for xyz, npoint, st in clouds:
        n = xyz.shape[1]
        start = torch.tensor([st])
        ref = stages.farthest_point_sample(xyz, npoint, start)
        cm = xyz.permute(0, 2, 1).contiguous().to(DEV)                             # [1, 3, n] channel-major
        index = F.build_index(lib.cloud_cm(cm), cm.device, 1, n)
        for mode in (2, 0):
            _, out = F.fps_indexed(lib.cloud_cm(cm), cm.device, 1, n, npoint, start, index, concurrent=mode)
            assert torch.equal(out.cpu().long(), ref), "mode %d, n %d, npoint %d" % (mode, n, npoint)


def test_fps_many_clouds_one_cta_each(dv):
    """More clouds than clusters fit: the library falls back to one CTA per cloud (batched rounds)."""
    g = torch.Generator().manual_seed(77)
    xyz = torch.rand(40, 4096, 3, generator=g) * 6 - 3
    start = torch.randint(0, 4096, (40,), generator=g)
    ref = stages.farthest_point_sample(xyz[:3], 4096, start[:3])
    out = dv.farthest_point_sample(xyz.to(DEV), 4096, start).cpu()
    assert torch.equal(out[:3], ref)
    assert torch.equal(out.sort(dim=1)[0], torch.arange(4096).expand(40, -1))


def test_fps_plain_and_pruned_kernels_agree(dv, F):
    g = torch.Generator().manual_seed(3)
    xyz = (torch.randn(3, 3000, 3, generator=g) * 5).to(DEV)
    start = torch.tensor([0, 17, 2999])
    assert torch.equal(F.fps_plain(xyz, 3000, start), dv.farthest_point_sample(xyz, 3000, start))


def test_fps_float64_and_padding(dv, prim):
    xyz = T(prim["xyz"])
    ref = T(prim["fps_f64"])
    out = dv.farthest_point_sample(xyz.double().to(DEV), 300, ref[:, 0])
    assert torch.equal(out.cpu(), ref)
    ref = T(prim["fps_pad"])
    out = dv.farthest_point_sample(xyz[:, :10].contiguous().to(DEV), 16, ref[:, 0])
    assert torch.equal(out.cpu(), ref)


def test_fps_matches_reference_fixture(dv, prim):
    for tag, key in (("f32", "xyz"), ("lat", "xyz_l")):
        ref = T(prim["fps_" + tag])
        out = dv.farthest_point_sample(T(prim[key]).to(DEV), 300, ref[:, 0])
        assert torch.equal(out.cpu(), ref)


def test_fps_default_start_follows_cpu_rng(dv):
    xyz = torch.rand(2, 500, 3).to(DEV)
    torch.manual_seed(123)
    expect = torch.randint(0, 500, (2,), dtype=torch.long)
    torch.manual_seed(123)
    out = dv.farthest_point_sample(xyz, 8)
    assert torch.equal(out[:, 0].cpu(), expect)


# ------------------------------------------------- square distance / ball ----
def test_square_distance_bit_exact(dv, prim):
    out = dv.square_distance(T(prim["q"]).to(DEV), T(prim["xyz"]).to(DEV))
    assert torch.equal(out.cpu(), T(prim["sqd"]))


def test_ball_query_reference_fixture(dv, prim):
    out = dv.query_ball_point(0.2, 16, T(prim["xyz"]).to(DEV), T(prim["q"]).to(DEV))
    assert torch.equal(out.cpu(), T(prim["ball_r02_n16"]))
    xl = T(prim["xyz_l"]).to(DEV)
    out = dv.query_ball_point(0.25, 8, xl, xl[:, :40].contiguous())
    assert torch.equal(out.cpu(), T(prim["ball_l_r025_n8"]))


@pytest.mark.parametrize("n,s,radius,nsample", [(777, 100, 0.5, 32), (9000, 333, 1.0, 256), (16384, 64, 0.3, 8)])
def test_ball_query_vs_oracle(dv, n, s, radius, nsample):
    xyz = lattice_cloud(n, n, extent=6.0, step=0.25)
    q = xyz[:, torch.randperm(n, generator=torch.Generator().manual_seed(1))[:s]].contiguous()
    ref = stages.query_ball_point(radius, nsample, xyz, q)
    out = dv.query_ball_point(radius, nsample, xyz.to(DEV), q.to(DEV))
    assert torch.equal(out.cpu(), ref)


def test_ball_query_empty_ball_yields_n(dv):
    xyz = torch.zeros(1, 40, 3)
    q = torch.full((1, 2, 3), 100.0)
    out = dv.query_ball_point(0.1, 4, xyz.to(DEV), q.to(DEV))
    assert (out == 40).all()
    with pytest.raises(IndexError):
        dv.query_ball_point(0.1, 64, xyz.to(DEV), q.to(DEV))   # N < nsample (reference :106)


def test_index_points_and_sample_and_group(dv, prim):
    out = dv.index_points(T(prim["xyz"]).to(DEV), T(prim["ip_idx"]).to(DEV))
    assert torch.equal(out.cpu(), T(prim["ip_out"]))
    xyz = T(prim["xyz"])
    new_xyz_ref = T(prim["sag_new_xyz"])
    start = torch.stack([(xyz[b] == new_xyz_ref[b, 0]).all(dim=1).nonzero()[0, 0] for b in range(2)])
    nx, npts, idx = dv.sample_and_group(32, 0.4, 8, xyz.to(DEV), T(prim["sag_feats"]).to(DEV), returnidx=True,
                                        start=start)
    assert torch.equal(idx.cpu(), T(prim["sag_idx"]))
    assert torch.equal(nx.cpu(), new_xyz_ref)
    assert torch.equal(npts.cpu(), T(prim["sag_new_points"]))


# ------------------------------------------------------ SA layer / WL --------
@pytest.mark.parametrize("name", ["fwd_modelnet_n1024_g5", "fwd_kitti_n2048_g7"])
def test_feat_extraction_vs_reference_fixture(dv, name):
    g = load_golden(name)
    sd = golden_state_dict(g)
    use_normal = g["src"].shape[1] == 6
    N = g["src"].shape[2]
    fe = dv.feat_extraction_layer(use_normal=use_normal, npoint=N)
    fe.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith("FE1.")})
    fe = fe.to(DEV).eval()
    xyz, feats = fe(T(g["src"]).to(DEV), start=torch.tensor([int(g["starts"][0])]))
    ref = T(g["src_fe_feat"])
    assert feats.shape == ref.shape
    assert rel_err(feats, ref) < FEAT_RTOL
    assert rel_err(feats, ref) < 1e-5          # in practice float32 round-off only
    src_xyz = T(g["src"])[:, :3].permute(0, 2, 1)
    assert torch.equal(xyz.cpu(), stages.index_points(src_xyz.contiguous(), T(g["src_fps"]).long()))


def test_sa_layer_dense_ball(dv):
    """Many members per ball (beyond nsample) and a feature tail: oracle comparison."""
    g = torch.Generator().manual_seed(5)
    N = 600
    pts = torch.cat([torch.rand(2, 3, N, generator=g), torch.randn(2, 3, N, generator=g)], dim=1)
    fe = dv.feat_extraction_layer(use_normal=True, npoint=N, radius=0.3, nsample=16)
    for bn in fe.sa1.mlp_bns:
        bn.running_mean.normal_(0, 0.1, generator=g)
        bn.running_var.uniform_(0.5, 1.5, generator=g)
    sd = {"FE1." + k: v for k, v in fe.state_dict().items()}
    start = torch.tensor([3, 77])
    _, ref, _ = stages.feat_extraction(sd, pts, start, radius=0.3, nsample=16)
    fe = fe.to(DEV).eval()
    _, out = fe(pts.to(DEV), start=start)
    assert rel_err(out, ref) < 1e-5


@pytest.mark.parametrize("kind,n", [("kitti", 16384), ("kitti", 5000), ("modelnet", 1024)])
def test_sa_pruned_kernel_equals_bruteforce_kernel(dv, F, synthetic, kind, n):
    src, _, _, _ = synthetic.make_batch(kind, [3, 4], n)
    fe = dv.feat_extraction_layer(use_normal=kind == "modelnet", npoint=n).to(DEV).eval()
    start = torch.tensor([1, 2])
    try:
        F.USE_INDEX = True
        _, a = fe(src.to(DEV), start=start)
        F.USE_INDEX = False
        _, b = fe(src.to(DEV), start=start)
    finally:
        F.USE_INDEX = True
    assert torch.equal(a, b)


@pytest.mark.parametrize("radius,nsample", [(0.5, 16), (0.12, 64)])
def test_sa_pruned_crowded_balls(dv, radius, nsample):
    """More members than nsample (smallest indices kept) and more than the kernel's
    member list (brute-force overflow pass): both against the oracle."""
    g = torch.Generator().manual_seed(8)
    N = 3000
    pts = torch.cat([torch.rand(1, 3, N, generator=g) * 0.6, torch.randn(1, 3, N, generator=g)], dim=1)
    fe = dv.feat_extraction_layer(use_normal=True, npoint=N, radius=radius, nsample=nsample)
    sd = {"FE1." + k: v for k, v in fe.state_dict().items()}
    start = torch.tensor([9])
    _, ref, _ = stages.feat_extraction(sd, pts, start, radius=radius, nsample=nsample)
    _, out = fe.to(DEV).eval()(pts.to(DEV), start=start)
    assert rel_err(out, ref) < 1e-5


def test_weighting_topk(dv, prim):
    sd = golden_state_dict(prim, "wl_sd/")
    wl = dv.weighting_layer()
    wl.load_state_dict({k[3:]: v for k, v in sd.items()})
    wl = wl.to(DEV)
    x = T(prim["wl_x"]).to(DEV)
    scores = wl.scores(x)
    ref_scores = stages.weighting_scores(sd, T(prim["wl_x"]))
    assert rel_err(scores, ref_scores) < 1e-6
    assert torch.equal(wl(x).cpu(), T(prim["wl_out"]))


def test_topk_ties_lowest_index_first(F):
    s = torch.tensor([[1.0, 3.0, 3.0, 2.0, 3.0, 0.5, 2.0, 7.0] + [0.0] * 100]).to(DEV)
    assert F.topk(s, 6).cpu().tolist() == [[7, 1, 2, 4, 3, 6]]


# ------------------------------------------------------ candidates / KNN -----
@pytest.mark.parametrize("G", [5, 6, 7, 11, 15])
def test_candidates_bit_exact(dv, prim, G):
    r = float(prim["vox_r%d" % G])
    out = dv.voxelize(T(prim["vox_centres"]).to(DEV), r, 0.4)
    assert torch.equal(out.cpu(), T(prim["vox_G%d" % G]))


@pytest.mark.parametrize("n,q,k", [(100, 37, 1), (1000, 500, 5), (5000, 1000, 32), (16384, 2000, 32), (8193, 65, 32)])
def test_knn_vs_oracle_with_ties(dv, n, q, k):
    ref_pts = lattice_cloud(n, n + 1, extent=10.0, step=0.5)          # exact distance ties
    g = torch.Generator().manual_seed(2)
    qry = torch.round((torch.rand(1, q, 3, generator=g) * 2 - 1) * 40) / 4
    d_ref, i_ref = stages.knn(ref_pts, qry, k)
    d, i = dv.KNN(k, transpose_mode=True)(ref_pts.to(DEV), qry.to(DEV))
    assert torch.equal(i.cpu(), i_ref)
    assert torch.equal(d.cpu(), d_ref)
    d2, i2 = dv.KNN(k, transpose_mode=False)(ref_pts.transpose(1, 2).to(DEV), qry.transpose(1, 2).to(DEV))
    assert torch.equal(i2.cpu(), i_ref.transpose(1, 2))


@pytest.mark.parametrize("n,q,k,chain", [(100, 37, 1, 1), (1000, 500, 5, 7), (5000, 1000, 32, 11), (16384, 3000, 32, 11),
                                          (8193, 65, 32, 3), (70, 300, 32, 11)])
def test_knn_indexed_vs_oracle_with_ties(F, n, q, k, chain):
    """Spatially pruned KNN == brute-force KNN == oracle, bit for bit (lattice clouds: exact ties)."""
    lib = importlib.import_module(PKG + "._lib")
    ref_pts = lattice_cloud(n, n + 1, extent=10.0, step=0.5)
    g = torch.Generator().manual_seed(2)
    qry = torch.round((torch.rand(1, q, 3, generator=g) * 2 - 1) * 60) / 4      # some queries far outside the cloud
    d_ref, i_ref = stages.knn(ref_pts, qry, k)
    rp = ref_pts.to(DEV)
    index = F.build_index(lib.cloud_pm(rp), rp.device, 1, n)
    d, i, i32 = F.knn_indexed(index, 0, rp.device, 1, n, qry.to(DEV), k, chain=chain, want32=True)
    assert torch.equal(i.cpu(), i_ref)
    d3, i3, _ = F.knn_indexed(index, 0, rp.device, 1, n, qry.to(DEV), k, chain=chain * 4, zline=chain)
    assert torch.equal(i3.cpu(), i_ref) and torch.equal(d3.cpu(), d_ref)
    assert torch.equal(i32.cpu().long(), i_ref)
    assert torch.equal(d.cpu(), d_ref)


def test_spatial_index_is_a_permutation_with_tight_boxes(F):
    lib = importlib.import_module(PKG + "._lib")
    n = 5000
    pts = lattice_cloud(n, 3).to(DEV)
    index = F.build_index(lib.cloud_pm(pts), pts.device, 1, n)
    spt = index.sorted_pt[0].cpu()                        # [cap, 4]: x, y, z, index bits
    sid = spt[:, 3].contiguous().view(torch.int32)
    valid = sid >= 0
    assert int(valid.sum()) == n and torch.equal(sid[valid].sort()[0], torch.arange(n, dtype=torch.int32))
    assert bool(torch.isinf(spt[~valid, :3]).all())       # unused slots are unreachable
    sxyz = spt[:, :3].T                                   # [3, cap]
    assert torch.equal(sxyz[:, valid].T, pts[0].cpu()[sid[valid].long()])
    box = index.bucket_box[0].cpu()
    for j in (0, 7, index.cap // 32 - 1):
        sl = slice(j * 32, j * 32 + 32)
        v = valid[sl]
        assert int(box[j, 6]) == int(v.sum())
        if v.any():
            p = sxyz[:, sl][:, v]
            assert torch.equal(box[j, :3], p.min(dim=1)[0]) and torch.equal(box[j, 3:6], p.max(dim=1)[0])


@pytest.mark.parametrize("n", [16385, 20000, 40000, 65536, 70000, 131072])
def test_big_spatial_index_is_a_permutation_with_tight_boxes(F, synthetic, n):
    """Multi-CTA index build (csrc/index_big.cu) for clouds above 16384 points."""
    lib = importlib.import_module(PKG + "._lib")
    pts = lattice_cloud(n, n, extent=40.0, step=0.125).to(DEV)       # duplicates and equal Morton codes
    index = F.build_index(lib.cloud_pm(pts), pts.device, 1, n, big=True)
    assert index.cap >= n and index.cap in (32768, 65536, 131072)
    spt = index.sorted_pt[0].cpu()
    sid = spt[:, 3].contiguous().view(torch.int32)
    valid = sid >= 0
    assert int(valid.sum()) == n and torch.equal(sid[valid].sort()[0], torch.arange(n, dtype=torch.int32))
    assert bool(valid[:n].all())                           # unused slots sort behind every point
    assert bool(torch.isinf(spt[~valid, :3]).all())
    assert torch.equal(spt[valid, :3], pts[0].cpu()[sid[valid].long()])
    box = index.bucket_box[0].cpu()
    cnt = valid.view(-1, 32).sum(dim=1)
    assert torch.equal(box[:, 6].long(), cnt)
    used = cnt > 0
    p = spt[:, :3].view(-1, 32, 3)
    lo = torch.where(valid.view(-1, 32, 1), p, torch.full_like(p, float("inf"))).min(dim=1)[0]
    hi = torch.where(valid.view(-1, 32, 1), p, torch.full_like(p, float("-inf"))).max(dim=1)[0]
    assert torch.equal(box[used, :3], lo[used]) and torch.equal(box[used, 3:6], hi[used])
    # Morton order: buckets are compact (ideal cubic cell of 32 points: edge 80 m * (32 / n)^(1/3))
    diag = (hi[used] - lo[used]).norm(dim=1)
    assert float(diag.median()) < 3.0 * 80.0 * (32.0 / n) ** (1.0 / 3.0)


@pytest.mark.parametrize("n,kind", [(20000, "kitti"), (40000, "kitti"), (65536, "lattice"), (70000, "kitti"),
                                     (131072, "kitti"), (131072, "lattice")])
def test_knn_big_index_equals_brute_force_and_oracle(F, n, kind):
    """KNN through the multi-CTA index (caps 32768 / 65536 / 131072, the last with the 17-bit key layout):
    every query equal to the brute-force kernel, a sample equal to the oracle; voxelised clouds: exact ties."""
    lib = importlib.import_module(PKG + "._lib")
    g = torch.Generator().manual_seed(n)
    if kind == "kitti":
        rho = (torch.randn(n, generator=g) * 25.0).abs().clamp(max=80.0)
        az = torch.rand(n, generator=g) * 6.2831853
        z = torch.rand(n, generator=g) * 8.0 - 2.0
        pts = (torch.round(torch.stack([rho * az.cos(), rho * az.sin(), z], 1) * 10.0) / 10.0).float().unsqueeze(0)
    else:
        pts = lattice_cloud(n, n + 1, extent=12.0, step=0.5)             # heavy duplicates: index tie-breaking
    centres = pts[0, torch.randint(0, n, (12,), generator=g)].double().unsqueeze(0)
    cand = F.candidates(centres.to(DEV), 1.2, 0.4, 7).view(1, -1, 3)      # 12 x 7^3 lattice queries
    far = (torch.rand(1, 64, 3, generator=g) * 2 - 1) * 150.0             # and some far outside the cloud
    qry = torch.cat([cand, far.to(DEV)], 1).contiguous()
    pd = pts.to(DEV)
    index = F.build_index(lib.cloud_pm(pd), pd.device, 1, n, big=True)
    d, i, i32 = F.knn_indexed(index, 0, pd.device, 1, n, qry, 32, chain=49, zline=7, want32=True)
    db, ib, _ = F.knn(lib.cloud_pm(pd), pd.device, 1, n, qry, 32)
    assert torch.equal(i, ib) and torch.equal(d, db) and torch.equal(i32.long(), ib)
    d1, i1, _ = F.knn_indexed(index, 0, pd.device, 1, n, qry, 5, chain=1)   # every query a cold start
    assert torch.equal(i1, ib[..., :5]) and torch.equal(d1, db[..., :5])
    pick = torch.arange(0, qry.shape[1], 23)
    d_ref, i_ref = stages.knn(pts, qry[:, pick].cpu(), 32)
    assert torch.equal(i[:, pick].cpu(), i_ref) and torch.equal(d[:, pick].cpu(), d_ref)


def test_knn_module_uses_big_index_same_results(dv, F):
    """knn_cuda.KNN stand-in above 16384 points (indexed path) == brute-force kernel."""
    lib = importlib.import_module(PKG + "._lib")
    n = 50000
    g = torch.Generator().manual_seed(5)
    pts = (torch.round(torch.randn(2, n, 3, generator=g) * 200.0) / 10.0).to(DEV)
    qry = (torch.round(torch.randn(2, 700, 3, generator=g) * 200.0) / 10.0).to(DEV)
    d, i = dv.KNN(32, transpose_mode=True)(pts, qry)
    db, ib, _ = F.knn(lib.cloud_pm(pts), pts.device, 2, n, qry, 32)
    assert torch.equal(i, ib) and torch.equal(d, db)


def test_knn_kitti_full_size_sampled_and_sorted(dv, F, synthetic):
    _, tgt, _, _ = synthetic.make_batch("kitti", [2], 16384)
    g = torch.Generator().manual_seed(4)
    centres = (torch.rand(1, 64, 3, generator=g, dtype=torch.float64) * 2 - 1) * 30
    cand = F.candidates(centres.to(DEV), 2.0, 0.4).view(1, -1, 3)      # 85184 queries
    tg = tgt.to(DEV)
    from importlib import import_module
    lib = import_module(PKG + "._lib")
    d, i, _ = F.knn(lib.cloud_cm(tg), tg.device, 1, 16384, cand, 32)
    index = F.build_index(lib.cloud_cm(tg), tg.device, 1, 16384)
    d2, i2, _ = F.knn_indexed(index, 0, tg.device, 1, 16384, cand, 32, chain=121, zline=11)
    assert torch.equal(i, i2) and torch.equal(d, d2)      # all 85184 queries: pruned == brute force
    d, i = d.cpu(), i.cpu()
    assert (d[..., 1:] >= d[..., :-1]).all()
    assert int(i.min()) >= 0 and int(i.max()) < 16384
    pick = torch.randperm(cand.shape[1], generator=g)[:1500]
    d_ref, i_ref = stages.knn(tgt[:, :3].permute(0, 2, 1).contiguous(), cand[:, pick].cpu(), 32)
    assert torch.equal(i[:, pick], i_ref)
    assert torch.equal(d[:, pick], d_ref)


# ------------------------------------------------------------- DFE / CPG -----
def test_dfe_dense_vs_reference_fixture(dv, prim):
    sd = golden_state_dict(prim, "dfe_sd/")
    dfe = dv.feat_embedding_layer()
    dfe.load_state_dict({k[4:]: v for k, v in sd.items()})
    dfe = dfe.to(DEV)
    out = dfe(T(prim["dfe_xs"]).to(DEV), src=True)
    assert out.shape == prim["dfe_src_out"].shape
    assert rel_err(out, T(prim["dfe_src_out"])) < 1e-5
    out = dfe(T(prim["dfe_xt"]).to(DEV), src=False)
    assert out.shape == prim["dfe_tgt_out"].shape
    assert rel_err(out, T(prim["dfe_tgt_out"])) < 1e-5


def test_cat_feat_tgt_module_and_fused_dfe(dv, F):
    g = load_golden("fwd_modelnet_n1024_g5")
    sd = golden_state_dict(g)
    cand = T(g["candidates"])
    tgt = T(g["tgt"])
    txyz = tgt[:, :3].permute(0, 2, 1).contiguous()
    tfeat = T(g["tgt_fe_feat"])
    s = int(g["stride"])
    # standalone module returns the reference's float64 tensor
    cat = dv.Get_Cat_Feat_Tgt()(cand.to(DEV), None, txyz.to(DEV), tfeat.to(DEV))
    assert cat.dtype == torch.float64
    ref = T(g["tgt_cat_s"])
    assert torch.allclose(cat[:, ::8, ::s].cpu(), ref, rtol=0, atol=1e-6)
    # fused gather + DFE against the reference's DFE output
    lib = importlib.import_module(PKG + "._lib")
    dfe = dv.feat_embedding_layer()
    dfe.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith("DFE.")})
    dfe = dfe.to(DEV)
    B, M, C, _ = cand.shape
    tg = tgt.to(DEV)
    cq = cand.view(B, M * C, 3).to(DEV)
    kd, _, ki = F.knn(lib.cloud_cm(tg), tg.device, B, 1024, cq, 32, want64=False, want32=True)
    out = F.dfe_tgt_fused(cq, lib.cloud_cm(tg), tfeat.to(DEV), kd, ki, B, 1024, dfe.params(), lib.QUIRKS_REFERENCE)
    out = out.view(B, M, C, 32)[:, :, ::s]
    assert rel_err(out, T(g["tgt_dfe_s"])) < 1e-5


def test_dfe_tensor_core_kernel_vs_fp32_kernel(dv, F, synthetic):
    """tcgen05 / TMEM embedding (collapsed map, 3xTF32) against the FP32 CUDA-core kernel
    and the oracle: within 1e-3 relative (north star), in practice ~1e-6."""
    lib = importlib.import_module(PKG + "._lib")
    g = load_golden("fwd_modelnet_n1024_g5")
    sd = golden_state_dict(g)
    dfe = dv.feat_embedding_layer()
    dfe.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith("DFE.")})
    dfe = dfe.to(DEV)
    cand = T(g["candidates"])
    B, M, C, _ = cand.shape
    tg = T(g["tgt"]).to(DEV)
    tfeat = T(g["tgt_fe_feat"]).to(DEV)
    cq = cand.view(B, M * C, 3).to(DEV)
    kd, _, ki = F.knn(lib.cloud_cm(tg), tg.device, B, 1024, cq, 32, want64=False, want32=True)
    ref = F.dfe_tgt_fused(cq, lib.cloud_cm(tg), tfeat, kd, ki, B, 1024, dfe.params(), lib.QUIRKS_REFERENCE)
    b_hi, b_lo = dfe.tc_operand()
    for nq in (M * C, 4 * 7 + 1, 3):                   # full, ragged tile tail, less than one tile
        out = F.dfe_tgt_tc(cq[:, :nq].contiguous(), lib.cloud_cm(tg), tfeat, kd[:, :nq].contiguous(),
                           ki[:, :nq].contiguous(), B, 1024, b_hi, b_lo, lib.QUIRKS_REFERENCE)
        torch.cuda.synchronize()
        assert rel_err(out, ref[:, :nq]) < FEAT_RTOL
        assert rel_err(out, ref[:, :nq]) < 2e-5
    s = int(g["stride"])
    out = F.dfe_tgt_tc(cq, lib.cloud_cm(tg), tfeat, kd, ki, B, 1024, b_hi, b_lo, lib.QUIRKS_REFERENCE)
    assert rel_err(out.view(B, M, C, 32)[:, :, ::s], T(g["tgt_dfe_s"])) < 2e-5


def test_cpg_standalone_vs_reference_fixture(dv, prim):
    sd = golden_state_dict(prim, "cpg_sd/")
    net = dv.cpg()
    net.load_state_dict({k[4:]: v for k, v in sd.items()})
    net = net.to(DEV)
    a, b, c = T(prim["cpg_a"]).to(DEV), T(prim["cpg_b"]).to(DEV), T(prim["cpg_c"]).to(DEV)
    out = net(a, b, c, 1, 0.4)                                   # contiguous [B,N,32,C]: layout 0
    assert torch.allclose(out.cpu(), T(prim["cpg_out"]), rtol=0, atol=5e-6)
    bt = b.permute(0, 1, 3, 2).contiguous().permute(0, 1, 3, 2)   # the view deepVCP.py:106 produces: layout 1
    out = net(a, bt, c, 1, 0.4)
    assert torch.allclose(out.cpu(), T(prim["cpg_out"]), rtol=0, atol=5e-6)


# --------------------------------------------------------------- Kabsch ------
def test_kabsch_vs_reference_fixture(dv, prim):
    x, y = T(prim["kab_x"]), T(prim["kab_y"])
    R, t = dv.get_rigid_transform(x.to(DEV), y.to(DEV))
    assert R.dtype == torch.float64
    assert torch.allclose(R.cpu(), T(prim["kab_R"]), atol=1e-12)
    assert torch.allclose(t.cpu(), T(prim["kab_t"]), atol=1e-12)
    R32, t32 = dv.get_rigid_transform(x.float().to(DEV), y.float().to(DEV))
    assert R32.dtype == torch.float32
    assert rot_angle_deg(R32, T(prim["kab_R"])) < 1e-3


def test_kabsch_reflection_not_corrected(dv):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 3, 50, generator=g, dtype=torch.float64)
    y = x.clone()
    y[:, 2] *= -1                                                   # mirrored target
    R, _ = dv.get_rigid_transform(x.to(DEV), y.to(DEV))
    Rr, _ = stages.get_rigid_transform(x, y)
    assert torch.allclose(R.cpu(), Rr, atol=1e-10)
    assert torch.det(R[0].cpu()) < 0                                # quirk Q10


def test_svd_optimization_vs_reference_fixture(dv, prim):
    R2, t2, _, _ = dv.svd_optimization(T(prim["kab_x"]).to(DEV), T(prim["kab_y"]).float().to(DEV),
                                       T(prim["svdopt_Rt"]).to(DEV), T(prim["svdopt_tt"]).to(DEV))
    assert torch.allclose(R2.cpu(), T(prim["svdopt_R2"]), atol=1e-10)
    assert torch.allclose(t2.cpu(), T(prim["svdopt_t2"]), atol=1e-10)


def test_kabsch_sweep_batch(dv, F):
    g = torch.Generator().manual_seed(99)
    B = 4096
    x = torch.randn(B, 3, 64, generator=g)
    Rg = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))[0]
    Rg = Rg * torch.sign(torch.det(Rg)).view(B, 1, 1)
    y = Rg @ x + torch.randn(B, 3, 1, generator=g) + 0.01 * torch.randn(B, 3, 64, generator=g)
    R, t = F.kabsch(x.to(DEV), y.to(DEV))
    Rr, tr = stages.get_rigid_transform(x.double(), y.double())
    assert rot_angle_deg(R, Rr) < ROT_TOL_DEG
    assert (t.cpu() - tr).abs().max() < TRANS_TOL


# -------------------------------------------------------- whole forward ------
def build_model(dv, g, N, tensor_cores=True):
    use_normal = g["src"].shape[1] == 6
    model = dv.DeepVCP(use_normal=use_normal, npoint=N, r=float(g["r"]), s=float(g["s"]))
    model.load_state_dict(golden_state_dict(g))
    model.dfe_tensor_cores = tensor_cores
    return model.to(DEV).eval()


def topk_equivalent(scores, a, b):
    if not torch.equal(scores[a], scores[b]):
        return False
    return True


@pytest.mark.parametrize("tensor_cores", [True, False])
@pytest.mark.parametrize("name", ["fwd_modelnet_n1024_g5", "fwd_modelnet_n512_g6", "fwd_kitti_n2048_g7"])
def test_forward_vs_reference_fixture(dv, name, tensor_cores):
    g = load_golden(name)
    N = int(g["n_points"])
    model = build_model(dv, g, N, tensor_cores)
    st = g["starts"]
    starts = tuple(torch.tensor([int(v)]) for v in st)
    src, tgt, R = T(g["src"]), T(g["tgt"]), T(g["R"])
    ref_topk = T(g["topk_idx"]).long().view(1, -1)
    # pass 1: free-running, checks the index-producing stages and the key-point choice
    model(src.to(DEV), tgt.to(DEV), R.to(DEV), torch.zeros(1, 3), starts=starts, keep_stages=True)
    L = model.last
    assert torch.equal(L["src_fps"].cpu(), T(g["src_fps"]))
    assert torch.equal(L["tgt_fps"].cpu(), T(g["tgt_fps"]))
    assert rel_err(L["src_fe_feat"], T(g["src_fe_feat"])) < 1e-5
    sc = L["scores"][0].cpu()
    mine = L["topk_idx"][0].cpu()
    # the reference's choice and ours may differ only inside groups of (near-)equal scores
    assert torch.allclose(sc[mine], sc[ref_topk[0]], rtol=1e-6, atol=0)
    # pass 2: teacher-forced key-point choice (torch.topk tie order is unspecified)
    kp, vcp = model(src.to(DEV), tgt.to(DEV), R.to(DEV), torch.zeros(1, 3), starts=starts, keep_stages=True,
                    topk_override=ref_topk)
    L = model.last
    assert torch.equal(L["src_keypts_full"].cpu(), T(g["src_keypts_full"]))
    assert torch.equal(L["picked_idx"].cpu().to(torch.int16), T(g["picked_idx"]))
    assert torch.allclose(L["src_cat"].cpu(), T(g["src_cat"]), atol=1e-6)
    assert torch.allclose(L["centres"].cpu(), T(g["centres"]), rtol=0, atol=1e-12)
    assert rel_err(L["src_dfe"], T(g["src_dfe"]).squeeze(2)) < 1e-5
    s = int(g["stride"])
    assert rel_err(L["tgt_dfe"][:, :, ::s], T(g["tgt_dfe_s"])) < FEAT_RTOL
    assert rel_err(L["tgt_dfe"][:, :, ::s], T(g["tgt_dfe_s"])) < 1e-5
    assert torch.equal(kp.cpu(), T(g["src_keypts"]))
    assert (vcp.cpu() - T(g["vcp"])).abs().max() < 2e-5
    # candidates: float64 centres may differ in the last bit -> compare with 1 ulp slack
    assert torch.allclose(L["candidates"].cpu(), T(g["candidates"]), rtol=0, atol=4e-6)
    # pose (train.py:110): two-stage Kabsch against the reference's own result
    R2, t2 = dv.pose_from_forward(kp, vcp, R.to(DEV), T(g["t"]).view(1, 3, 1).to(DEV))
    assert rot_angle_deg(R2, T(g["R2"])) < ROT_TOL_DEG
    assert (t2.cpu() - T(g["t2"])).abs().max() < TRANS_TOL


def test_forward_knn_indices_vs_oracle(dv):
    """KNN indices inside the forward are bit-exact against the oracle on the same candidates."""
    g = load_golden("fwd_kitti_n2048_g7")
    model = build_model(dv, g, 2048)
    starts = tuple(torch.tensor([int(v)]) for v in g["starts"])
    ref_topk = T(g["topk_idx"]).long().view(1, -1)
    model(T(g["src"]).to(DEV), T(g["tgt"]).to(DEV), T(g["R"]).to(DEV), torch.zeros(1, 3), starts=starts,
          keep_stages=True, topk_override=ref_topk)
    L = model.last
    cand = L["candidates"].cpu()
    txyz = T(g["tgt"])[:, :3].permute(0, 2, 1).contiguous()
    d_ref, i_ref = stages.knn(txyz, cand.view(1, -1, 3), 32)
    assert torch.equal(L["knn_idx"].cpu(), i_ref)
    assert torch.equal(L["knn_dist"].cpu(), d_ref)


def test_batched_forward_equals_independent_forwards(dv, synthetic):
    N = 1024
    src, tgt, R, t = synthetic.make_batch("modelnet", [10, 11, 12], N)
    torch.manual_seed(0)
    model = dv.DeepVCP(use_normal=True, npoint=N, r=0.8, s=0.4).to(DEV).eval()
    starts = (torch.tensor([1, 2, 3]), torch.tensor([4, 5, 6]), torch.tensor([7, 8, 9]))
    kp, vcp = model(src.to(DEV), tgt.to(DEV), R.to(DEV), torch.zeros(1, 3), starts=starts)
    for b in range(3):
        st = tuple(s[b:b + 1] for s in starts)
        kb, vb = model(src[b:b + 1].to(DEV), tgt[b:b + 1].to(DEV), R[b:b + 1].to(DEV), torch.zeros(1, 3), starts=st)
        assert torch.equal(kb[0], kp[b])
        assert torch.equal(vb[0], vcp[b])


def test_forward_seeded_rng_matches_explicit_starts(dv, synthetic):
    N = 1024
    src, tgt, R, _ = synthetic.make_batch("modelnet", [20], N)
    torch.manual_seed(0)
    model = dv.DeepVCP(use_normal=True, npoint=N, r=0.8, s=0.4).to(DEV).eval()
    torch.manual_seed(42)
    a = torch.randint(0, N, (1,)), torch.randint(0, 64, (1,)), torch.randint(0, N, (1,))
    torch.manual_seed(42)
    kp1, v1 = model(src.to(DEV), tgt.to(DEV), R.to(DEV), torch.zeros(1, 3))
    kp2, v2 = model(src.to(DEV), tgt.to(DEV), R.to(DEV), torch.zeros(1, 3), starts=a)
    assert torch.equal(kp1, kp2) and torch.equal(v1, v2)


def test_cpu_tensors_to_ops_raise(dv):
    with pytest.raises(RuntimeError):
        dv.query_ball_point(0.2, 4, torch.rand(1, 10, 3), torch.rand(1, 2, 3))
    with pytest.raises(RuntimeError):
        dv.DeepVCP(use_normal=True).eval()(torch.rand(1, 6, 64), torch.rand(1, 6, 64),
                                           torch.eye(3, dtype=torch.float64)[None], torch.zeros(1, 3))


def test_kitti_shaped_forward_full_size_properties(dv, synthetic):
    """BASELINE config K8 shape, B=2: index-stage properties + oracle on the cheap stages."""
    N = 16384
    src, tgt, R, t = synthetic.make_batch("kitti", [0, 1], N)
    torch.manual_seed(1)
    model = dv.DeepVCP(use_normal=False, npoint=N, r=2.0, s=0.4).to(DEV).eval()
    starts = (torch.tensor([11, 12]), torch.tensor([13, 14]), torch.tensor([15, 16]))
    kp, vcp = model(src.to(DEV), tgt.to(DEV), R.to(DEV), torch.zeros(1, 3), starts=starts, keep_stages=True)
    L = model.last
    assert kp.shape == (2, 64, 3) and vcp.shape == (2, 64, 3)
    assert torch.isfinite(vcp).all()
    for key in ("src_fps", "tgt_fps"):
        assert torch.equal(L[key].cpu().long().sort(dim=1)[0], torch.arange(N).expand(2, N))
    d = L["knn_dist"].cpu()
    assert (d[..., 1:] >= d[..., :-1]).all()
    # vcp is a convex combination of the candidates of its key-point
    cand = L["candidates"].cpu()
    assert (vcp.cpu() >= cand.min(dim=2)[0] - 1e-4).all() and (vcp.cpu() <= cand.max(dim=2)[0] + 1e-4).all()
    # FPS of the full-size cloud against the oracle (pair 0, src)
    ref = stages.farthest_point_sample(src[:1, :3].permute(0, 2, 1).contiguous(), N, starts[0][:1])
    assert torch.equal(L["src_fps"][:1].cpu().long(), ref)
    R2, t2 = dv.pose_from_forward(kp, vcp, R.to(DEV), t.view(2, 3, 1).to(DEV))
    assert torch.isfinite(R2).all() and torch.isfinite(t2).all()


def test_overlapped_sa_equals_sequential_path(dv, F, synthetic):
    """K8-shaped forward: the SA layer that runs beside the sampling (original point order, rows
    gathered into FPS order afterwards) must give the features of the plain FPS -> SA sequence."""
    lib = importlib.import_module(PKG + "._lib")
    N = 16384
    src, tgt, R, _ = synthetic.make_batch("kitti", [3], N)
    torch.manual_seed(5)
    model = dv.DeepVCP(use_normal=False, npoint=N, r=2.0, s=0.4).to(DEV).eval()
    starts = (torch.tensor([7]), torch.tensor([8]), torch.tensor([9]))
    model(src.to(DEV), tgt.to(DEV), R.to(DEV), torch.zeros(1, 3), starts=starts, keep_stages=True)
    L = model.last
    both = torch.cat([src, tgt], 0).to(DEV)
    st2 = torch.cat([starts[0], starts[2]])
    index = F.SpatialIndex(2, N, both.device)
    _, fps2 = F.fps(lib.cloud_cm(both), both.device, both.dtype, 2, N, N, st2, want64=False, want32=True, index=index)
    sa = model.FE1.sa1
    _, feat2 = F.sa_layer(lib.cloud_cm(both), None, 0, fps2, 2, N, N, sa.radius, sa.nsample, sa.folded(), both.device,
                          want_xyz=False, index=index)
    assert torch.equal(fps2[:1], L["src_fps"]) and torch.equal(fps2[1:], L["tgt_fps"])
    assert torch.equal(feat2[:1], L["src_fe_feat"]) and torch.equal(feat2[1:], L["tgt_fe_feat"])


@pytest.mark.parametrize("depth", [2, 3, 4])
def test_streamed_registration_equals_one_batch_at_a_time(dv, synthetic, depth):
    """Throughput mode (batches alternating between streams; from depth 3 on several feature halves in flight
    and one sampling CTA per cloud) returns, in order, exactly the poses of the same batches registered one
    after the other."""
    N = 4096
    torch.manual_seed(9)
    model = dv.DeepVCP(use_normal=False, npoint=N, r=1.2, s=0.4).to(DEV).eval()
    batches = []
    for i in range(7):
        src, tgt, R, t = synthetic.make_batch("kitti", [2 * i, 2 * i + 1], N)
        starts = (torch.tensor([i, i + 1]), torch.tensor([i + 2, i + 3]), torch.tensor([i + 4, i + 5]))
        batches.append((src, tgt, R, t.view(2, 3, 1), starts))
    ref = []
    for src, tgt, R, t, starts in batches:
        kp, vcp = model(src.to(DEV), tgt.to(DEV), R.to(DEV), torch.zeros(1, 3), starts=starts)
        R2, t2 = dv.pose_from_forward(kp, vcp, R.to(DEV), t.to(DEV))
        ref.append(dv.sharding.pack_poses(R2, t2).cpu())
    pipe = dv.StreamedRegistration(model, depth=depth)
    assert pipe.sampling == (2 if depth >= 3 else 0) and len(pipe.fe_streams) == max(1, depth - 1)
    hosts = [torch.empty(2, 12, dtype=torch.float64).pin_memory() for _ in batches]
    for (src, tgt, R, t, starts), h in zip(batches, hosts):
        pipe.submit(src, tgt, R, R, t, starts=starts, host_out=h)
    out = pipe.collect()
    for a, b, h in zip(out, ref, hosts):
        assert torch.equal(a.cpu(), b) and torch.equal(h, b)


def test_graphed_registration_equals_streamed(dv, synthetic):
    """The forward + pose solve captured into CUDA graphs (GraphedRegistration) and replayed for a stream of
    different batches gives bit-identical poses to the eager streamed path, in submission order."""
    N, B = 4096, 2
    torch.manual_seed(5)
    model = dv.DeepVCP(use_normal=False, npoint=N, r=1.2000000000000002, s=0.4).to(DEV).eval()
    batches = []
    for k in range(9):
        src, tgt, R, t = synthetic.make_batch("kitti", [2 * k, 2 * k + 1], N)
        starts = (torch.tensor([k, k + 1]), torch.tensor([k + 2, k + 3]), torch.tensor([k + 4, k + 5]))
        batches.append((src, tgt, R, t.view(B, 3, 1), starts))
    eager = dv.StreamedRegistration(model, depth=2)
    for src, tgt, R, t, starts in batches:
        eager.submit(src.to(DEV), tgt.to(DEV), R.to(DEV), R.to(DEV), t.to(DEV), starts=starts)
    ref = [p.cpu() for p in eager.collect()]
    for depth in (3, 4, 2, 1):   # 3, 4: one sampling CTA per cloud, 2 / 3 feature halves in flight
        gr = dv.GraphedRegistration(model, B, 3, N, depth=depth)
        assert gr.launches_per_batch >= 10
        host = [torch.empty(B, 12, dtype=torch.float64).pin_memory() for _ in batches]
        for i, (src, tgt, R, t, starts) in enumerate(batches):
            if i % 2:    # device inputs, device outputs
                gr.submit(src.to(DEV), tgt.to(DEV), R.to(DEV), R.to(DEV), t.to(DEV), starts)
            else:        # pinned host inputs, pinned host outputs
                gr.submit(src.pin_memory(), tgt.pin_memory(), R.pin_memory(), R.pin_memory(), t.pin_memory(), starts,
                          host_out=host[i])
        out = gr.collect()
        for i, p in enumerate(out):
            assert torch.equal(p.cpu(), ref[i]), "batch %d differs (depth %d)" % (i, depth)


def test_reference_native_shape_with_normals_vs_oracle(dv, synthetic):
    """The reference's own operating point (deepVCP.py:76-77, deep_feat_extraction.py:10): N = 10000
    points WITH normals, r = 1.0, s = 0.4 (6^3 candidates). Exercises the overlapped feature half
    (cluster FPS, by-bucket SA layer with 6 input channels, row gather) against the CPU oracle."""
    N = 10000
    src, tgt, R, t = synthetic.make_batch("modelnet", [21], N)
    src[:, :3] *= 8.0                       # unit ball -> 8 m: balls of 0.1 m hold a handful of points
    tgt[:, :3] *= 8.0
    torch.manual_seed(3)
    model = dv.DeepVCP(use_normal=True).eval()           # all reference literals: npoint 10000, r 1.0, s 0.4
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    starts = (torch.tensor([17]), torch.tensor([5]), torch.tensor([4242]))
    ref = stages.deepvcp_forward(sd, src, tgt, R, 1.0, 0.4, starts)
    model = model.to(DEV)
    kp, vcp = model(src.to(DEV), tgt.to(DEV), R.to(DEV), torch.zeros(1, 3), starts=starts, keep_stages=True,
                    topk_override=ref["topk_idx"])
    L = model.last
    assert torch.equal(L["src_fps"].cpu().long(), ref["src_fps"]) and torch.equal(L["tgt_fps"].cpu().long(), ref["tgt_fps"])
    assert rel_err(L["src_fe_feat"].cpu(), ref["src_fe_feat"]) < 1e-5
    assert rel_err(L["tgt_fe_feat"].cpu(), ref["tgt_fe_feat"]) < 1e-5
    assert torch.equal(kp.cpu(), ref["src_keypts"])
    assert torch.equal(L["candidates"].cpu(), ref["candidates"])
    assert torch.equal(L["knn_idx"].cpu(), ref["knn_idx"])
    assert (vcp.cpu() - ref["vcp"]).abs().max() < 1e-4
    R2, t2 = dv.pose_from_forward(kp, vcp, R.to(DEV), t.view(1, 3, 1).to(DEV))
    R2r, t2r, _, _, _ = stages.pose_from_forward(ref["src_keypts"], ref["vcp"], R, t.view(1, 3, 1))
    assert rot_angle_deg(R2, R2r) < ROT_TOL_DEG and (t2.cpu() - t2r).abs().max() < TRANS_TOL


def test_forward_vs_reference_record_at_native_operating_point(dv, synthetic):
    """CUDA path against the record of the UNMODIFIED reference with all its literals as shipped
    (tests/golden/make_golden.py:native_case: N = 10000 with normals, radius 0.1 / nsample 256,
    r = 1.0, s = 0.4 -> 6^3 candidates). Indices bit-exact, features 1e-5, pose within the north-star bar."""
    g = load_golden("fwd_reference_native_n10000_g6")
    N = int(g["n_points"])
    src, tgt, R, t = synthetic.make_batch(str(g["kind"]), [int(g["pair_id"])], N)
    model = dv.DeepVCP(use_normal=True)                   # reference literals
    model.load_state_dict(golden_state_dict(g))
    model = model.to(DEV).eval()
    starts = tuple(torch.tensor([int(v)]) for v in g["starts"])
    ref_topk = T(g["topk_idx"]).long().view(1, -1)
    model(src.to(DEV), tgt.to(DEV), R.to(DEV), torch.zeros(1, 3), starts=starts, keep_stages=True)
    sc = model.last["scores"][0].cpu()
    assert torch.allclose(sc[model.last["topk_idx"][0].cpu()], sc[ref_topk[0]], rtol=1e-6, atol=0)
    kp, vcp = model(src.to(DEV), tgt.to(DEV), R.to(DEV), torch.zeros(1, 3), starts=starts, keep_stages=True,
                    topk_override=ref_topk)
    L = model.last
    assert torch.equal(L["src_fps"].cpu().to(torch.int16), T(g["src_fps"]))
    assert torch.equal(L["tgt_fps"].cpu().to(torch.int16), T(g["tgt_fps"]))
    assert torch.equal(L["picked_idx"].cpu().to(torch.int16), T(g["picked_idx"]))
    assert torch.equal(L["src_keypts_full"].cpu(), T(g["src_keypts_full"]))
    s = int(g["stride"])
    assert rel_err(L["src_fe_feat"][:, ::s], T(g["src_fe_feat_s"])) < 1e-5
    assert rel_err(L["tgt_fe_feat"][:, ::s], T(g["tgt_fe_feat_s"])) < 1e-5
    assert torch.allclose(L["centres"].cpu(), T(g["centres"]), rtol=0, atol=1e-12)
    assert torch.allclose(L["candidates"].cpu().view(1, 64, -1, 3)[:, ::4], T(g["candidates_s"]), rtol=0, atol=4e-6)
    assert rel_err(L["src_dfe"], T(g["src_dfe"]).squeeze(2)) < 1e-5
    assert rel_err(L["tgt_dfe"][:, :, ::s], T(g["tgt_dfe_s"])) < 1e-5
    assert torch.equal(kp.cpu(), T(g["src_keypts"]))
    assert (vcp.cpu() - T(g["vcp"])).abs().max() < 2e-5
    R2, t2 = dv.pose_from_forward(kp, vcp, R.to(DEV), t.view(1, 3, 1).to(DEV))
    assert rot_angle_deg(R2, T(g["R2"])) < ROT_TOL_DEG
    assert (t2.cpu() - T(g["t2"])).abs().max() < TRANS_TOL


# ---------------------------------------- intended-semantics mode ------
@pytest.mark.parametrize("quirks", [0, 63 - 4, 63 - 8, 63 - 1, 63 - 2, 63 - 16, 63 - 32, 32, 4 + 16])
def test_forward_with_quirk_switches_vs_oracle(dv, synthetic, quirks):
    """Every quirk bit cleared on its own and all of them cleared ("intended" mode: proper key-point
    permute, per-neighbour weights, un-scrambled cost volume, t_init applied, reflection fix, feature rows
    addressed in their own order -- Q3, Q7, Q4, Q6, Q10, Q5): CUDA
    path against the oracle with the same switches. Indices bit-exact, features 1e-5, pose north-star."""
    N = 1024
    src, tgt, R, t = synthetic.make_batch("modelnet", [31, 32], N)
    torch.manual_seed(11)
    model = dv.DeepVCP(use_normal=True, npoint=N, r=0.8, s=0.4, quirks=quirks).eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    starts = (torch.tensor([5, 6]), torch.tensor([7, 8]), torch.tensor([9, 10]))
    t_init = torch.tensor([[0.25, -0.5, 0.125], [-1.0, 0.0, 0.75]])
    ref = stages.deepvcp_forward(sd, src, tgt, R, 0.8, 0.4, starts, quirks=quirks, t_init=t_init)
    model = model.to(DEV)
    kp, vcp = model(src.to(DEV), tgt.to(DEV), R.to(DEV), t_init, starts=starts, keep_stages=True,
                    topk_override=ref["topk_idx"])
    L = model.last
    assert torch.equal(L["src_keypts_full"].cpu(), ref["src_keypts_full"])
    assert torch.equal(L["picked_idx"].cpu(), ref["picked_idx"])
    assert torch.allclose(L["centres"].cpu(), ref["centres"], rtol=0, atol=1e-12)
    if not (quirks & 8):
        assert not torch.allclose(L["centres"].cpu(), (R @ ref["src_keypts"].transpose(1, 2).double()).transpose(1, 2))
    assert torch.allclose(L["candidates"].cpu(), ref["candidates"], rtol=0, atol=4e-6)
    if torch.equal(L["candidates"].cpu(), ref["candidates"]):
        assert torch.equal(L["knn_idx"].cpu(), ref["knn_idx"])
    assert rel_err(L["src_dfe"], ref["src_dfe"]) < 1e-5
    assert rel_err(L["tgt_dfe"], ref["tgt_dfe"]) < 1e-5
    assert rel_err(L["logits"], ref["logits"].view_as(L["logits"])) < 1e-4
    assert (vcp.cpu() - ref["vcp"]).abs().max() < 5e-5
    R2, t2 = dv.pose_from_forward(kp, vcp, R.to(DEV), t.view(2, 3, 1).to(DEV), quirks=quirks)
    R2r, t2r, _, _, _ = stages.pose_from_forward(ref["src_keypts"], ref["vcp"], R, t.view(2, 3, 1), quirks=quirks)
    assert rot_angle_deg(R2, R2r) < ROT_TOL_DEG and (t2.cpu() - t2r).abs().max() < TRANS_TOL
    if not (quirks & 16):
        assert (torch.det(R2.cpu()) > 0).all()


def test_kabsch_reflection_fix_and_weights_vs_oracle(dv, F):
    g = torch.Generator().manual_seed(17)
    B, n = 512, 64
    x = torch.randn(B, 3, n, generator=g, dtype=torch.float64)
    Rg = torch.linalg.qr(torch.randn(B, 3, 3, generator=g, dtype=torch.float64))[0]
    Rg = Rg * torch.sign(torch.det(Rg)).view(B, 1, 1)
    Rg[1::2, :, 0] = -Rg[1::2, :, 0]                                                       # every other one improper
    y = Rg @ x + torch.randn(B, 3, 1, generator=g, dtype=torch.float64) + 0.01 * torch.randn(B, 3, n, generator=g, dtype=torch.float64)
    w = torch.rand(B, n, generator=g, dtype=torch.float64)
    w[:, ::7] = 0.0
    for quirks, fix in ((dv.QUIRKS_REFERENCE, False), (0, True)):
        for weights in (None, w):
            R, t = F.kabsch(x.to(DEV), y.to(DEV), quirks=quirks, weights=None if weights is None else weights.to(DEV))
            Rr, tr = stages.get_rigid_transform(x, y, reflection_fix=fix, weights=weights)
            assert rot_angle_deg(R, Rr) < ROT_TOL_DEG and (R.cpu() - Rr).abs().max() < 1e-9
            assert (t.cpu() - tr).abs().max() < 1e-9
            if fix:
                assert torch.allclose(torch.det(R.cpu()), torch.ones(B, dtype=torch.float64), atol=1e-9)
    Rq, _ = F.kabsch(x.to(DEV), y.to(DEV))
    assert (torch.det(Rq.cpu()) < 0).any() and (torch.det(Rq.cpu()) > 0).any()          # reference mode keeps reflections
    R2, t2, R1, t1 = F.kabsch_refine(x.to(DEV), y.to(DEV), Rg.to(DEV), torch.zeros(B, 3, 1, dtype=torch.float64).to(DEV),
                                     want_first=True, quirks=0)
    R2r, t2r, R1r, t1r, _ = stages.svd_optimization(x, y, Rg, torch.zeros(B, 3, 1, dtype=torch.float64), reflection_fix=True)
    assert (R1.cpu() - R1r).abs().max() < 1e-9 and (R2.cpu() - R2r).abs().max() < 1e-8


# ------------------------------------------------------------ data ingest ------
def test_kitti_ingest_matches_the_reference_loader_arithmetic(dv, synthetic):
    """KITTIDataset.py:11-16,44-46,67-84 restated with numpy (seeded np.random like the reference)."""
    rs = np.random.RandomState(5)
    scans = [rs.randn(m, 4).astype(np.float32) * 20 for m in (30000, 23456, 16500)]
    N = 16384
    idx = np.stack([dv.KITTIDataset.downsample_indices(s.shape[0], N, rs) for s in scans])
    _, _, R, t = synthetic.make_batch("kitti", [0, 1, 2], 64)
    src, tgt, refl = dv.KITTIDataset.ingest(scans, idx, R, t, want_reflectance=True)
    for b, s in enumerate(scans):
        pts = s[idx[b], :]                                   # downsample
        sp = pts[:, :3].T                                    # 3 x N float32
        assert np.array_equal(src[b].cpu().numpy(), sp)
        assert np.array_equal(refl[b, 0].cpu().numpy(), pts[:, 3])
        tg = (R[b].numpy() @ sp + t[b].numpy().reshape(3, 1)).astype(np.float32)      # float64 like numpy promotes
        assert np.allclose(tgt[b].cpu().numpy(), tg, rtol=2e-7, atol=1e-6)
    # the ingested pair goes straight into the model
    model = dv.DeepVCP(use_normal=False, npoint=N, r=2.0, s=0.4).to(DEV).eval()
    kp, vcp = model(src[:1], tgt[:1], R[:1].to(DEV), torch.zeros(1, 3))
    assert torch.isfinite(vcp).all()
    with pytest.raises(IndexError):
        dv.KITTIDataset.ingest(scans, np.full((3, 8), 29999), R, t)


def test_pose_from_forward_kernel_equals_refine_on_permuted_copies(dv, F):
    """dvcp_pose_from_forward reads the forward's float32 [B,n,3] outputs in place; same arithmetic as
    dvcp_kabsch_refine on the permuted float64 copies the reference makes (deepVCP_loss.py:105-107)."""
    g = torch.Generator().manual_seed(23)
    B, n = 37, 64
    kp = torch.randn(B, n, 3, generator=g)
    Rg = torch.linalg.qr(torch.randn(B, 3, 3, generator=g, dtype=torch.float64))[0]
    tg = torch.randn(B, 3, 1, generator=g, dtype=torch.float64)
    vcp = ((Rg @ kp.double().transpose(1, 2) + tg).transpose(1, 2) + 0.05 * torch.randn(B, n, 3, generator=g)).float()
    for quirks in (dv.QUIRKS_REFERENCE, 0):
        R2, t2 = dv.pose_from_forward(kp.to(DEV), vcp.to(DEV), Rg.to(DEV), tg.to(DEV), quirks=quirks)
        Rr, tr, _, _ = F.kabsch_refine(kp.to(DEV).permute(0, 2, 1).double(), vcp.to(DEV).permute(0, 2, 1).double(),
                                       Rg.to(DEV), tg.to(DEV), quirks=quirks)
        assert torch.equal(R2, Rr) and torch.equal(t2, tr)
    kp6 = torch.randn(B, n, 6, generator=g).to(DEV)                     # use_normal=True: a strided slice
    R2, t2 = dv.pose_from_forward(kp6[:, :, :3], vcp.to(DEV), Rg.to(DEV), tg.to(DEV))
    Rr, tr, _, _ = F.kabsch_refine(kp6[:, :, :3].permute(0, 2, 1).double(), vcp.to(DEV).permute(0, 2, 1).double(),
                                   Rg.to(DEV), tg.to(DEV))
    assert torch.equal(R2, Rr) and torch.equal(t2, tr)


# ------------------------------------ repaired three-layer feature extraction ------
def test_chained_feature_extraction_and_forward_vs_oracle(dv, F, synthetic):
    """chained_fe=True: sa1 -> sa2 -> sa3 -> fc with xyz and features handed on (what
    deep_feat_extraction.py:10-15,26-28 intends; the reference itself crashes there).
    FPS indices of every layer bit-exact, features 1e-5, pose within the north-star bar."""
    N = 512
    src, tgt, R, t = synthetic.make_batch("modelnet", [41, 42], N)
    torch.manual_seed(13)
    model = dv.DeepVCP(use_normal=True, npoint=N, r=0.8, s=0.4, chained_fe=True).eval()
    assert model.FE1.sa2.mlp_convs[0].in_channels == 35 and model.FE1.sa3.mlp_convs[0].in_channels == 67
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    starts = (torch.randint(0, N, (3, 2), generator=g), torch.tensor([7, 8]), torch.randint(0, N, (3, 2), generator=g))
    ref = stages.deepvcp_forward(sd, src, tgt, R, 0.8, 0.4, starts, chained_fe=True)
    model = model.to(DEV)
    xyz3, feat3, fps3 = model.FE1(src.to(DEV), start=starts[0], return_fps=True)
    assert torch.equal(fps3.cpu().long(), ref["src_fps"])
    assert torch.equal(xyz3.cpu(), ref["src_fe_xyz"])
    assert rel_err(feat3, ref["src_fe_feat"]) < 1e-5
    kp, vcp = model(src.to(DEV), tgt.to(DEV), R.to(DEV), torch.zeros(1, 3), starts=starts, keep_stages=True,
                    topk_override=ref["topk_idx"])
    L = model.last
    assert torch.equal(L["tgt_fps"].cpu().long(), ref["tgt_fps"])
    assert rel_err(L["tgt_fe_feat"], ref["tgt_fe_feat"]) < 1e-5
    assert torch.equal(kp.cpu(), ref["src_keypts"])
    assert rel_err(L["tgt_dfe"], ref["tgt_dfe"]) < 1e-5
    assert (vcp.cpu() - ref["vcp"]).abs().max() < 5e-5
    R2, t2 = dv.pose_from_forward(kp, vcp, R.to(DEV), t.view(2, 3, 1).to(DEV))
    R2r, t2r, _, _, _ = stages.pose_from_forward(ref["src_keypts"], ref["vcp"], R, t.view(2, 3, 1))
    assert rot_angle_deg(R2, R2r) < ROT_TOL_DEG and (t2.cpu() - t2r).abs().max() < TRANS_TOL
    # the seeded draw order: three per cloud, the key-point draw in between
    torch.manual_seed(99)
    a = model.draw_starts(2, N)
    assert a[0].shape == (3, 2) and a[1].shape == (2,) and a[2].shape == (3, 2)


def test_dfe_tensor_core_kernel_ragged_batches(dv, F, synthetic):
    """Several clouds with a candidate count that is no multiple of a tile (4), a run (32) or anything else: runs of
    tiles cross cloud boundaries (the cloud index of a candidate comes from a float quotient + correction), the last
    run is clipped (the bulk copies of the index stream too). Against the FP32 CUDA-core kernel, both weight modes."""
    lib = importlib.import_module(PKG + "._lib")
    torch.manual_seed(21)
    dfe = dv.feat_embedding_layer().to(DEV)
    for B, N, Q in ((3, 1024, 125 * 7 + 3), (5, 300, 37), (2, 4096, 32 * 41 + 31)):
        g = torch.Generator().manual_seed(B * 1000 + Q)
        tg = (torch.rand(B, 3, N, generator=g) * 4 - 2).to(DEV)
        tfeat = torch.randn(B, N, 32, generator=g).to(DEV)
        cq = (torch.rand(B, Q, 3, generator=g) * 4 - 2).to(DEV)
        kd, _, ki = F.knn(lib.cloud_cm(tg), tg.device, B, N, cq, 32, want64=False, want32=True)
        b_hi, b_lo = dfe.tc_operand()
        for quirks in (lib.QUIRKS_REFERENCE, lib.QUIRKS_INTENDED):
            ref = F.dfe_tgt_fused(cq, lib.cloud_cm(tg), tfeat, kd, ki, B, N, dfe.params(), quirks)
            out = F.dfe_tgt_tc(cq, lib.cloud_cm(tg), tfeat, kd, ki, B, N, b_hi, b_lo, quirks)
            assert rel_err(out, ref) < 2e-5, (B, N, Q, quirks)
