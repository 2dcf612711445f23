"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so parity is pinned
on outputs of the reference itself: this script imports it under the shims of
oracle/reference_shims.py, feeds it the seeded synthetic inputs of
deepvcp-pointcloud-registration_b200/synthetic.py and stores inputs, weights and
every stage-boundary tensor. Large tensors are stored as a strided sample plus
an order-sensitive checksum. The KNN inside the reference forward is the
stand-in for the absent third-party `knn_cuda` (parity unpinned for that stage,
see oracle/__init__.py).
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import reference_shims as rs  # noqa: E402

syn = importlib.import_module("deepvcp-pointcloud-registration_b200.synthetic")


def checksum(a: np.ndarray) -> np.int64:
    """Order-sensitive 64-bit checksum of an integer array (wrap-around)."""
    v = np.ascontiguousarray(a).astype(np.int64).reshape(-1)
    w = (np.arange(v.size, dtype=np.int64) * np.int64(2654435761)) ^ np.int64(0x9E3779B97F4A7C15 - (1 << 64))
    with np.errstate(over="ignore"):
        return np.int64(np.sum((v + 1) * (w | 1), dtype=np.int64))


def npy(t):
    return t.detach().cpu().numpy()


def forward_case(name, kind, n_points, r, s, model_seed, rng_seed, pair_id=0, stride=16, dtype="f32"):
    use_normal = kind == "modelnet"
    src, tgt, R, t = syn.make_batch(kind, [pair_id], n_points, dtype=dtype)
    model = rs.make_model(use_normal, n_points, seed=model_seed)
    m = rs.load()
    pu = m.pointnet2_utils
    fps_log, ball_log = [], []
    orig_fps, orig_ball = pu.farthest_point_sample, pu.query_ball_point

    def fps(xyz, npoint):
        out = orig_fps(xyz, npoint)
        fps_log.append(out.clone())
        return out

    def ball(radius, nsample, xyz, new_xyz):
        out = orig_ball(radius, nsample, xyz, new_xyz)
        ball_log.append(out.clone())
        return out

    pu.farthest_point_sample, pu.query_ball_point = fps, ball
    rec = {}
    try:
        torch.manual_seed(rng_seed)
        kp, vcp = rs.forward(model, src, tgt, R, torch.zeros(1, 3), r, s, rec)
    finally:
        pu.farthest_point_sample, pu.query_ball_point = orig_fps, orig_ball
    assert len(fps_log) == 3 and len(ball_log) == 3
    # pose solve exactly as train.py:110 drives it
    with rs.quiet():
        x = kp.permute(0, 2, 1).double()
        y = vcp.permute(0, 2, 1).double()
        R1, t1 = m.deepVCP_loss.get_rigid_transform(x, y)
        R2, t2, x1, y2 = m.deepVCP_loss.svd_optimization(x, y, R, t.view(1, 3, 1))
    out = {
        "kind": kind, "r": r, "s": s, "n_points": n_points, "dtype": dtype, "pair_id": pair_id,
        "src": npy(src), "tgt": npy(tgt), "R": npy(R), "t": npy(t),
        "starts": np.array([int(fps_log[0][0, 0]), int(fps_log[1][0, 0]), int(fps_log[2][0, 0])]),
        "src_fps": npy(fps_log[0]).astype(np.int32), "kp_fps": npy(fps_log[1]).astype(np.int32),
        "tgt_fps": npy(fps_log[2]).astype(np.int32),
        "src_ball": npy(ball_log[0]).astype(np.int32 if n_points > 32767 else np.int16),
        "kp_ball": npy(ball_log[1]).astype(np.int16),
        "src_fe_feat": npy(rec["src_fe_feat"]), "tgt_fe_feat": npy(rec["tgt_fe_feat"]),
        "topk_idx": npy(rec["topk_idx"]).astype(np.int32),
        "src_keypts_full": npy(rec["src_keypts_full"]),
        "picked_idx": npy(rec["picked_idx"]).astype(np.int16),
        "src_cat": npy(rec["src_cat"]),
        "centres": npy(rec["centres"]),
        "candidates": npy(rec["candidates"]),
        "src_dfe": npy(rec["src_dfe"]),
        "vcp": npy(vcp), "src_keypts": npy(kp),
        "R1": npy(R1), "t1": npy(t1), "R2": npy(R2), "t2": npy(t2), "x1": npy(x1), "y_pred2": npy(y2),
        "stride": stride,
    }
    # target-side tensors: strided sample over candidates (axis 2)
    tgt_dfe = rec["tgt_dfe"].permute(0, 1, 3, 2)                 # [1,64,C,32]
    out["tgt_dfe_s"] = npy(tgt_dfe[:, :, ::stride])
    tc = rec["tgt_cat"]                                          # [1,64,C,32,35] f64
    out["tgt_cat_s"] = npy(tc[:, ::8, ::stride]).astype(np.float64)
    for k, v in model.state_dict().items():
        out["sd/" + k] = npy(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def native_case(name="fwd_reference_native_n10000_g6", pair_id=5, model_seed=4, rng_seed=12, stride=16):
    """The reference at its OWN operating point: all literals as shipped (npoint = 10000, radius 0.1,
    nsample 256, r = 1.0, s = 0.4 -> 6^3 candidates, use_normal=True), no (r, s) shim. The clouds are
    regenerated from (kind, pair_id, n_points) by the tests; big tensors are stored as a strided sample
    plus an order-sensitive checksum."""
    n_points = 10000
    src, tgt, R, t = syn.make_batch("modelnet", [pair_id], n_points)
    model = rs.make_model(True, n_points, seed=model_seed)
    m = rs.load()
    pu = m.pointnet2_utils
    fps_log, ball_log = [], []
    orig_fps, orig_ball = pu.farthest_point_sample, pu.query_ball_point

    def fps(xyz, npoint):
        out = orig_fps(xyz, npoint)
        fps_log.append(out.clone())
        return out

    def ball(radius, nsample, xyz, new_xyz):
        out = orig_ball(radius, nsample, xyz, new_xyz)
        ball_log.append(out.clone())
        return out

    pu.farthest_point_sample, pu.query_ball_point = fps, ball
    rec = {}
    try:
        torch.manual_seed(rng_seed)
        kp, vcp = rs.forward(model, src, tgt, R, torch.zeros(1, 3), 1.0, 0.4, rec)
    finally:
        pu.farthest_point_sample, pu.query_ball_point = orig_fps, orig_ball
    with rs.quiet():
        x = kp.permute(0, 2, 1).double()
        y = vcp.permute(0, 2, 1).double()
        R2, t2, x1, y2 = m.deepVCP_loss.svd_optimization(x, y, R, t.view(1, 3, 1))
    out = {
        "kind": "modelnet", "pair_id": pair_id, "n_points": n_points, "r": 1.0, "s": 0.4, "stride": stride,
        "R": npy(R), "t": npy(t),
        "starts": np.array([int(fps_log[0][0, 0]), int(fps_log[1][0, 0]), int(fps_log[2][0, 0])]),
        "src_fps": npy(fps_log[0]).astype(np.int16), "tgt_fps": npy(fps_log[2]).astype(np.int16),
        "kp_fps": npy(fps_log[1]).astype(np.int16),
        "src_ball_checksum": checksum(npy(ball_log[0])), "tgt_ball_checksum": checksum(npy(ball_log[2])),
        "src_ball_s": npy(ball_log[0][:, ::250]).astype(np.int16),
        "src_fe_feat_s": npy(rec["src_fe_feat"][:, ::stride]), "tgt_fe_feat_s": npy(rec["tgt_fe_feat"][:, ::stride]),
        "topk_idx": npy(rec["topk_idx"]).astype(np.int32),
        "src_keypts_full": npy(rec["src_keypts_full"]),
        "picked_idx": npy(rec["picked_idx"]).astype(np.int16),
        "centres": npy(rec["centres"]),
        "candidates_s": npy(rec["candidates"][:, ::4]),
        "src_dfe": npy(rec["src_dfe"]),
        "tgt_dfe_s": npy(rec["tgt_dfe"].permute(0, 1, 3, 2)[:, :, ::stride]),
        "vcp": npy(vcp), "src_keypts": npy(kp), "R2": npy(R2), "t2": npy(t2),
    }
    for k, v in model.state_dict().items():
        out["sd/" + k] = npy(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def primitives_case():
    """Small calls of the L1 primitives and the standalone modules."""
    m = rs.load()
    pu = m.pointnet2_utils
    g = torch.Generator().manual_seed(11)
    out = {}
    xyz = torch.rand(2, 300, 3, generator=g) * 2 - 1
    # lattice-snapped copy: exact ties in distance
    xyz_l = torch.round(xyz * 8) / 8
    q = xyz[:, :40].contiguous()
    out["xyz"], out["xyz_l"], out["q"] = npy(xyz), npy(xyz_l), npy(q)
    out["sqd"] = npy(pu.square_distance(q, xyz))
    out["ball_r02_n16"] = npy(pu.query_ball_point(0.2, 16, xyz, q))
    out["ball_l_r025_n8"] = npy(pu.query_ball_point(0.25, 8, xyz_l, xyz_l[:, :40].contiguous()))
    for tag, cloud in (("f32", xyz), ("lat", xyz_l), ("f64", xyz.double())):
        torch.manual_seed(3)
        out["fps_" + tag] = npy(pu.farthest_point_sample(cloud, 300))
    torch.manual_seed(5)
    out["fps_pad"] = npy(pu.farthest_point_sample(xyz[:, :10].contiguous(), 16))
    idx = torch.randint(0, 300, (2, 7, 5), generator=g)
    out["ip_idx"] = npy(idx)
    out["ip_out"] = npy(pu.index_points(xyz, idx))
    torch.manual_seed(9)
    feats = torch.randn(2, 300, 4, generator=g)
    nx, npnts, gi = pu.sample_and_group(32, 0.4, 8, xyz, feats, returnidx=True)
    out["sag_feats"], out["sag_new_xyz"], out["sag_new_points"], out["sag_idx"] = (
        npy(feats), npy(nx), npy(npnts), npy(gi))
    # voxelize for several (r, s)
    centres = (torch.rand(2, 5, 3, generator=g, dtype=torch.float64) * 40 - 20)
    out["vox_centres"] = npy(centres)
    for G, r in ((5, 0.8), (6, 1.0), (7, 1.2000000000000002), (11, 2.0), (15, 2.8000000000000003)):
        out["vox_G%d" % G] = npy(m.voxelize.voxelize(centres, r, 0.4))
        out["vox_r%d" % G] = np.float64(r)
    # cpg standalone, cpg.py:62-79 call convention
    torch.manual_seed(0)
    net = m.cpg.cpg()
    a = torch.randn(2, 6, 1, 32, generator=g)
    b = torch.randn(2, 6, 32, 216, generator=g)
    c = torch.randn(2, 6, 216, 3, generator=g)
    with torch.no_grad():
        out["cpg_out"] = npy(net(a, b, c, 1, 0.4))
    out["cpg_a"], out["cpg_b"], out["cpg_c"] = npy(a), npy(b), npy(c)
    for k, v in net.state_dict().items():
        out["cpg_sd/cpg." + k] = npy(v)
    # DFE standalone
    torch.manual_seed(1)
    dfe = m.deep_feat_embedding.feat_embedding_layer()
    xs = torch.randn(2, 5, 32, 35, generator=g, dtype=torch.float64)
    xt = torch.randn(2, 5, 9, 32, 35, generator=g, dtype=torch.float64)
    with torch.no_grad():
        out["dfe_src_out"], out["dfe_tgt_out"] = npy(dfe(xs, src=True)), npy(dfe(xt, src=False))
    out["dfe_xs"], out["dfe_xt"] = npy(xs), npy(xt)
    for k, v in dfe.state_dict().items():
        out["dfe_sd/DFE." + k] = npy(v)
    # weighting layer standalone (tie-free random scores)
    torch.manual_seed(2)
    wl = m.weighting_layer.weighting_layer()
    xw = torch.randn(2, 500, 32, generator=g)
    with torch.no_grad():
        out["wl_out"] = npy(wl(xw))
    out["wl_x"] = npy(xw)
    for k, v in wl.state_dict().items():
        out["wl_sd/WL." + k] = npy(v)
    # Kabsch, both dtypes, incl. a mirrored target (det = -1)
    x = torch.randn(3, 3, 64, generator=g, dtype=torch.float64)
    Rg = torch.linalg.qr(torch.randn(3, 3, 3, generator=g, dtype=torch.float64))[0]
    y = Rg @ x + torch.randn(3, 3, 1, generator=g, dtype=torch.float64) + 0.01 * torch.randn(3, 3, 64, generator=g, dtype=torch.float64)
    Rk, tk = m.deepVCP_loss.get_rigid_transform(x, y)
    out["kab_x"], out["kab_y"], out["kab_R"], out["kab_t"] = npy(x), npy(y), npy(Rk), npy(tk)
    Rt = Rg
    tt = torch.randn(3, 3, 1, generator=g, dtype=torch.float64)
    with rs.quiet():
        R2, t2, x1, y2 = m.deepVCP_loss.svd_optimization(x, y.float(), Rt, tt)
    out["svdopt_Rt"], out["svdopt_tt"], out["svdopt_R2"], out["svdopt_t2"] = npy(Rt), npy(tt), npy(R2), npy(t2)
    path = os.path.join(HERE, "primitives.npz")
    np.savez_compressed(path, **out)
    print("primitives ->", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def dtype_cases():
    """The dtypes the reference's own loaders produce: float64 clouds (ModelNet40Dataset.py:38,92) and a
    float32 scan with a float64 target (KITTIDataset.py:84,97)."""
    forward_case("fwd_modelnet_f64_n1024_g5", "modelnet", 1024, 0.8, 0.4, model_seed=5, rng_seed=21, pair_id=7,
                 dtype="f64")
    forward_case("fwd_kitti_mixed_n2048_g7", "kitti", 2048, syn.grid_radius(7), 0.4, model_seed=6, rng_seed=22,
                 pair_id=4, stride=32, dtype="mixed")


def train_case(name, kind, n_points, r, s, model_seed, rng_seed, pair_id=0, dtype="f32", lr=0.001):
    """One training step of the UNMODIFIED reference exactly as train.py:93-123 drives it: model.train(),
    forward, deepVCP_loss(alpha=0.5), backward, Adam(lr).step(). Stored: the inputs, the initial state_dict,
    the indices the forward drew / chose (FPS starts, top-k), the loss, R / t, the gradient of every parameter
    that received one, and the state_dict after the step (BatchNorm running statistics included)."""
    use_normal = kind == "modelnet"
    src, tgt, R, t = syn.make_batch(kind, [pair_id], n_points, dtype=dtype)
    model = rs.make_model(use_normal, n_points, seed=model_seed)
    model.train()
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    m = rs.load()
    pu = m.pointnet2_utils
    fps_log = []
    orig_fps = pu.farthest_point_sample

    def fps(xyz, npoint):
        out = orig_fps(xyz, npoint)
        fps_log.append(out.clone())
        return out

    pu.farthest_point_sample = fps
    rec = {}
    optim = torch.optim.Adam(model.parameters(), lr=lr)
    try:
        torch.manual_seed(rng_seed)
        kp, vcp = rs.forward(model, src, tgt, R, torch.zeros(1, 3), r, s, rec, grad=True)
    finally:
        pu.farthest_point_sample = orig_fps
    optim.zero_grad()
    with rs.quiet():
        loss, Rp, tp = m.deepVCP_loss.deepVCP_loss(kp, vcp, R, t.view(1, 3, 1), alpha=0.5)
    loss.backward()
    out = {
        "kind": kind, "r": r, "s": s, "n_points": n_points, "dtype": dtype, "pair_id": pair_id, "lr": lr,
        "src": npy(src), "tgt": npy(tgt), "R": npy(R), "t": npy(t),
        "starts": np.array([int(fps_log[0][0, 0]), int(fps_log[1][0, 0]), int(fps_log[2][0, 0])]),
        "topk_idx": npy(rec["topk_idx"]).astype(np.int32),
        "src_fe_feat": npy(rec["src_fe_feat"]), "vcp": npy(vcp), "src_keypts": npy(kp),
        "loss": npy(loss), "R_pred": npy(Rp), "t_pred": npy(tp),
    }
    for k, p in model.named_parameters():
        if p.grad is not None:
            out["grad/" + k] = npy(p.grad)
    optim.step()
    for k, v in sd0.items():
        out["sd/" + k] = npy(v)
    for k, v in model.state_dict().items():
        out["sd_after/" + k] = npy(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024),
          "loss %.6f, %d parameters with a gradient" % (float(loss), sum(k.startswith("grad/") for k in out)))


def train_cases():
    train_case("train_modelnet_n512_g5", "modelnet", 512, 0.8, 0.4, model_seed=7, rng_seed=31, pair_id=2)
    train_case("train_modelnet_f64_n512_g5", "modelnet", 512, 0.8, 0.4, model_seed=8, rng_seed=32, pair_id=6, dtype="f64")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "native":
        native_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "dtypes":
        dtype_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "train":
        train_cases()
        sys.exit(0)
    primitives_case()
    native_case()
    forward_case("fwd_modelnet_n1024_g5", "modelnet", 1024, 0.8, 0.4, model_seed=0, rng_seed=7)
    forward_case("fwd_modelnet_n512_g6", "modelnet", 512, 1.0, 0.4, model_seed=1, rng_seed=8, pair_id=3)
    forward_case("fwd_kitti_n2048_g7", "kitti", 2048, syn.grid_radius(7), 0.4, model_seed=2, rng_seed=9,
                 pair_id=1, stride=32)
    dtype_cases()
    train_cases()
