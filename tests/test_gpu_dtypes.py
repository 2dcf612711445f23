"""The dtypes the reference's own loaders hand to the model: float64 clouds (ModelNet40Dataset.py:38,92) and a
float32 scan with a float64 target (KITTIDataset.py:84,97). The CUDA path follows torch's promotion rules
(double FPS distances / ball query / grouping, `.float()` before the shared MLP and the embedding) and is held
here to records of the UNMODIFIED reference (tests/golden/make_golden.py:dtype_cases) and to the oracle."""
import importlib

import pytest
import torch

from conftest import PKG, golden_state_dict, load_golden
from oracle import stages

pytestmark = pytest.mark.gpu
T = torch.from_numpy
DEV = "cuda"


@pytest.fixture(scope="module")
def dv():
    return importlib.import_module(PKG)


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rot_angle_deg(Ra, Rb):
    d = Ra.double().cpu() @ Rb.double().cpu().transpose(-1, -2)
    c = ((d.diagonal(dim1=-2, dim2=-1).sum(-1) - 1) / 2).clamp(-1, 1)
    s = 0.5 * torch.stack([d[..., 2, 1] - d[..., 1, 2], d[..., 0, 2] - d[..., 2, 0], d[..., 1, 0] - d[..., 0, 1]], -1).norm(dim=-1)
    return torch.rad2deg(torch.atan2(s, c)).max().item()


@pytest.mark.parametrize("name", ["fwd_modelnet_f64_n1024_g5", "fwd_kitti_mixed_n2048_g7"])
def test_forward_float64_and_mixed_clouds_vs_reference_record(dv, name):
    g = load_golden(name)
    N = int(g["n_points"])
    src, tgt, R = T(g["src"]), T(g["tgt"]), T(g["R"])
    assert tgt.dtype == torch.float64 and src.dtype == (torch.float64 if "f64" in name else torch.float32)
    model = dv.DeepVCP(use_normal=src.shape[1] == 6, npoint=N, r=float(g["r"]), s=float(g["s"]))
    model.load_state_dict(golden_state_dict(g))
    model = model.to(DEV).eval()
    starts = tuple(torch.tensor([int(v)]) for v in g["starts"])
    ref_topk = T(g["topk_idx"]).long().view(1, -1)
    model(src.to(DEV), tgt.to(DEV), R.to(DEV), torch.zeros(1, 3), starts=starts, keep_stages=True)
    L = model.last
    assert torch.equal(L["src_fps"].cpu(), T(g["src_fps"])) and torch.equal(L["tgt_fps"].cpu(), T(g["tgt_fps"]))
    assert rel_err(L["src_fe_feat"], T(g["src_fe_feat"])) < 1e-5
    assert rel_err(L["tgt_fe_feat"], T(g["tgt_fe_feat"])) < 1e-5
    sc = L["scores"][0].cpu()
    assert torch.allclose(sc[L["topk_idx"][0].cpu()], sc[ref_topk[0]], rtol=1e-6, atol=0)
    kp, vcp = model(src.to(DEV), tgt.to(DEV), R.to(DEV), torch.zeros(1, 3), starts=starts, keep_stages=True,
                    topk_override=ref_topk)
    L = model.last
    assert kp.dtype == src.dtype and vcp.dtype == torch.float32          # deepVCP.py:110 dtypes
    assert torch.equal(L["src_keypts_full"].cpu(), T(g["src_keypts_full"]))
    assert torch.equal(L["picked_idx"].cpu().to(torch.int16), T(g["picked_idx"]))
    assert torch.allclose(L["src_cat"].cpu().double(), T(g["src_cat"]).double(), atol=1e-6)
    assert torch.allclose(L["centres"].cpu(), T(g["centres"]), rtol=0, atol=1e-12)
    assert torch.allclose(L["candidates"].cpu(), T(g["candidates"]), rtol=0, atol=4e-6)
    assert rel_err(L["src_dfe"], T(g["src_dfe"]).squeeze(2)) < 1e-5
    s = int(g["stride"])
    # the target's local coordinates are formed in float32 from the float32-rounded target (the reference subtracts
    # in double and rounds after): <= 1 ulp of a coordinate, inside the 1e-3 feature bar by orders of magnitude
    assert rel_err(L["tgt_dfe"][:, :, ::s], T(g["tgt_dfe_s"])) < 1e-4
    assert torch.equal(kp.cpu(), T(g["src_keypts"]))
    assert (vcp.cpu() - T(g["vcp"])).abs().max() < 5e-5
    t = T(g["t"]).view(1, 3, 1)
    R2, t2 = dv.pose_from_forward(kp, vcp, R.to(DEV), t.to(DEV))
    assert rot_angle_deg(R2, T(g["R2"])) < 1e-3 and (t2.cpu() - T(g["t2"])).abs().max() < 1e-4
    # svd_optimization returns the reference's x1 / y_pred2 (deepVCP_loss.py:81-90)
    x = kp.permute(0, 2, 1).double()
    y = vcp.permute(0, 2, 1).double()
    R2b, t2b, x1, y2 = dv.svd_optimization(x, y, R.to(DEV), t.to(DEV))
    assert x1.shape == g["x1"].shape and y2.shape == g["y_pred2"].shape
    # the inlier ORDER can differ from the record only where 1-NN distances are equal to the last bit
    assert torch.allclose(x1.cpu().sort(dim=2)[0], T(g["x1"]).sort(dim=2)[0], atol=1e-12)
    assert torch.allclose(y2.cpu().sort(dim=2)[0], T(g["y_pred2"]).sort(dim=2)[0], atol=1e-4)
    loss, Rl, tl = dv.deepVCP_loss(kp, vcp, R.to(DEV), t.to(DEV), 0.5)
    xo, yo = T(g["src_keypts"]), T(g["vcp"])
    R2r, t2r, _, _, inl = stages.pose_from_forward(xo, yo, R, t)
    x_in = torch.gather(xo.permute(0, 2, 1).double(), 2, inl.unsqueeze(1).expand(-1, 3, -1))
    y_opt = R2r @ x_in + t2r
    y_true = R @ x_in + t
    ref_loss = 0.5 * (y_true - y_opt).abs().mean() + 0.5 * (y_opt - y_true).mean().abs()
    assert abs(float(loss) - float(ref_loss)) < 1e-4 * max(1.0, float(ref_loss))


def test_float64_primitives_vs_oracle(dv):
    g = torch.Generator().manual_seed(12)
    xyz = torch.rand(2, 700, 3, generator=g, dtype=torch.float64) * 2 - 1
    lat = torch.round(xyz * 8) / 8                                      # exact ties at the ball boundary
    q = xyz[:, :90].contiguous()
    for cloud, qq, r, ns in ((xyz, q, 0.3, 16), (lat, lat[:, :90].contiguous(), 0.25, 8), (xyz, q, 0.05, 4)):
        ref = stages.query_ball_point(r, ns, cloud, qq)
        out = dv.query_ball_point(r, ns, cloud.to(DEV), qq.to(DEV))
        assert torch.equal(out.cpu(), ref)
    from oracle import native
    d = dv.square_distance(q.to(DEV), xyz.to(DEV))
    assert d.dtype == torch.float64 and torch.equal(d.cpu(), native.square_distance(q, xyz))
    # mixed: a float32 query against a float64 cloud promotes to double
    d2 = dv.square_distance(q.float().to(DEV), xyz.to(DEV))
    assert torch.equal(d2.cpu(), native.square_distance(q.float().double(), xyz))
    # sample_and_group in double, features stay float32
    feats = torch.randn(2, 700, 4, generator=g)
    start = torch.tensor([5, 9])
    nx, npnts, gi = dv.sample_and_group(64, 0.4, 8, xyz.to(DEV), feats.to(DEV), returnidx=True, start=start)
    rx, rp, ri, _ = stages.sample_and_group(64, 0.4, 8, xyz, feats, start)
    assert torch.equal(gi.cpu(), ri) and torch.equal(nx.cpu(), rx)
    assert npnts.dtype == torch.float64 and torch.allclose(npnts.cpu(), rp.double(), atol=0, rtol=0)


def test_set_abstraction_float64_cloud_vs_oracle(dv):
    """PointNetSetAbstraction.forward with float64 xyz + normals (pointnet2_utils.py:176-202): double ball query,
    double relative coordinates cast to float before the MLP."""
    g = torch.Generator().manual_seed(3)
    N = 900
    d = torch.randn(1, N, 3, generator=g, dtype=torch.float64)
    xyz = d / d.norm(dim=-1, keepdim=True) * torch.rand(1, N, 1, generator=g, dtype=torch.float64) ** (1 / 3)
    nrm = d / d.norm(dim=-1, keepdim=True)
    torch.manual_seed(0)
    sa = dv.PointNetSetAbstraction(N, 0.2, 32, 6, [16, 16, 32], False).eval()
    for bn in sa.mlp_bns:
        bn.running_mean.normal_(0, 0.1)
        bn.running_var.uniform_(0.75, 1.25)
    sd = {"FE1.sa1." + k: v.clone() for k, v in sa.state_dict().items()}
    start = torch.tensor([11])
    _, ref, _ = stages.set_abstraction(sd, "FE1.sa1", 3, xyz, nrm, N, 0.2, 32, start)
    sa = sa.to(DEV)
    new_xyz, feats = sa(xyz.permute(0, 2, 1).to(DEV), nrm.permute(0, 2, 1).to(DEV), start=start)
    assert new_xyz.dtype == torch.float64
    assert rel_err(feats.permute(0, 2, 1), ref) < 1e-5
