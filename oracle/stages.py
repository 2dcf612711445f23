"""CPU restatement of every stage of the DeepVCP registration hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). Index-producing stages call
the C restatement (oracle/dvcp_oracle.c through oracle/native.py); the dense
arithmetic is written with torch CPU ops in the dtype the reference uses at
that point. Every function cites the reference lines it restates (paths
relative to /root/reference) and SURVEY.md Appendix A/B for the pinned
semantics. Pinning: tests/test_oracle_vs_golden.py compares every stage with
tensors recorded from the reference itself (tests/golden/make_golden.py).

Batched semantics: B > 1 means B independent B = 1 forwards; the
reference itself only runs B = 1.
"""
import math

import torch
import torch.nn.functional as F

from . import native


# --------------------------------------------------------------------------- #
# L1 point-set primitives
# --------------------------------------------------------------------------- #
def farthest_point_sample(xyz, npoint, start):
    """pointnet2_utils.py:63-84 with the random start index made explicit."""
    return native.fps(xyz, npoint, start)


def query_ball_point(radius, nsample, xyz, new_xyz):
    """pointnet2_utils.py:87-107 (float32 clouds)."""
    return native.ball_query(radius, nsample, xyz, new_xyz)


def index_points(points, idx):
    """pointnet2_utils.py:43-60: points[b, idx[b, ...], :]."""
    B = points.shape[0]
    flat = idx.reshape(B, -1)
    out = torch.stack([points[b].index_select(0, flat[b]) for b in range(B)])
    return out.reshape(*idx.shape, points.shape[-1])


def sample_and_group(npoint, radius, nsample, xyz, points, start):
    """pointnet2_utils.py:110-138. Returns new_xyz, new_points, idx, fps_idx."""
    B, N, C = xyz.shape
    fps_idx = farthest_point_sample(xyz, npoint, start)
    new_xyz = index_points(xyz, fps_idx)
    idx = query_ball_point(radius, nsample, xyz, new_xyz)
    if int(idx.max()) >= N:
        raise IndexError("empty ball: index %d out of range" % int(idx.max()))
    grouped = index_points(xyz, idx) - new_xyz.view(B, npoint, 1, C)
    if points is not None:
        grouped = torch.cat([grouped, index_points(points, idx)], dim=-1)
    return new_xyz, grouped, idx, fps_idx


def _bn_eval(x, sd, prefix):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], False, 0.0, 1e-5)


def set_abstraction(sd, prefix, n_layers, xyz, points, npoint, radius, nsample, start):
    """PointNetSetAbstraction.forward, eval mode (pointnet2_utils.py:176-202).

    xyz [B,N,3], points [B,N,D] or None (point-major, unlike the reference's
    channel-major arguments). Returns new_xyz [B,S,3], feats [B,S,D'], fps_idx.
    Padding slots of the ball query repeat slot 0, so the max over nsample equals
    the max over the distinct members; the oracle still evaluates every slot.
    """
    new_xyz, grouped, idx, fps_idx = sample_and_group(npoint, radius, nsample, xyz, points, start)
    x = grouped.permute(0, 3, 2, 1).float()               # [B, 3+D, nsample, S]
    for i in range(n_layers):
        w = sd["%s.mlp_convs.%d.weight" % (prefix, i)]
        b = sd["%s.mlp_convs.%d.bias" % (prefix, i)]
        x = F.relu(_bn_eval(F.conv2d(x, w, b), sd, "%s.mlp_bns.%d" % (prefix, i)))
    feats = x.max(dim=2)[0].permute(0, 2, 1).contiguous()  # [B, S, D']
    return new_xyz, feats, fps_idx


def feat_extraction(sd, pts, start, npoint=None, radius=0.1, nsample=256, chained=False):
    """deep_feat_extraction.py:18-32, runnable semantics = sa1 only.

    pts [B,C_in,N] channel-major like the reference. Returns xyz [B,S,3],
    feats [B,S,32] in FPS order, fps_idx [B,S].
    chained=True (beyond the reference): the three-layer stack the file intends --
    sa1 -> sa2 (radius 0.2, 128 samples) -> sa3 (0.4, 64), xyz and features handed on, then fc (:10-15);
    start is then [3,B], one FPS start per layer; fps_idx is the last layer's.
    """
    xyz = pts[:, :3, :].permute(0, 2, 1).contiguous()
    nrm = pts[:, 3:, :].permute(0, 2, 1).contiguous() if pts.shape[1] > 3 else None
    npoint = pts.shape[2] if npoint is None else npoint
    if not chained:
        return set_abstraction(sd, "FE1.sa1", 3, xyz, nrm, npoint, radius, nsample, start)
    x1, f1, _ = set_abstraction(sd, "FE1.sa1", 3, xyz, nrm, npoint, radius, nsample, start[0])
    x2, f2, _ = set_abstraction(sd, "FE1.sa2", 2, x1, f1, npoint, 0.2, 128, start[1])
    x3, f3, fps3 = set_abstraction(sd, "FE1.sa3", 2, x2, f2, npoint, 0.4, 64, start[2])
    return x3, F.linear(f3, sd["FE1.fc.weight"], sd["FE1.fc.bias"]), fps3


# --------------------------------------------------------------------------- #
# weighting layer + top-K (weighting_layer.py:26-33)
# --------------------------------------------------------------------------- #
def weighting_scores(sd, feats):
    x = F.relu(F.linear(feats, sd["WL.fc1.0.weight"], sd["WL.fc1.0.bias"]))
    x = F.relu(F.linear(x, sd["WL.fc2.0.weight"], sd["WL.fc2.0.bias"]))
    x = F.softplus(F.linear(x, sd["WL.fc3.0.weight"], sd["WL.fc3.0.bias"]))
    return x.squeeze(-1)                                   # [B, S]


def topk_indices(scores, k=64):
    """Canonical order (score desc, index asc); torch.topk leaves ties
    unspecified."""
    B, S = scores.shape
    out = torch.empty(B, k, dtype=torch.int64)
    for b in range(B):
        order = sorted(range(S), key=lambda i: (-float(scores[b, i]), i))
        out[b] = torch.tensor(order[:k])
    return out


# --------------------------------------------------------------------------- #
# key-point stage (deepVCP.py:44-67)
# --------------------------------------------------------------------------- #
def gather_keypoints(src_pts, topk_idx, view_quirk=True):
    """deepVCP.py:44-46 incl. the view-instead-of-permute layout,
    with the channel count taken from the input (Q2). Per sample.
    view_quirk=False: the permute the code intends (key-point k = column k)."""
    B, C, N = src_pts.shape
    K = topk_idx.shape[1]
    out = torch.empty(B, K, C, dtype=src_pts.dtype)
    for b in range(B):
        g = src_pts[b][:, topk_idx[b]]                     # [C, K]
        out[b] = g.reshape(K, C) if view_quirk else g.t()  # row-major re-read / proper permute
    return out


def cat_feat_src(src_keypts, grouped, keyfeats):
    """get_cat_feat_src.py:37-53."""
    kp = src_keypts[:, :, :3].unsqueeze(2)
    diff = kp - grouped[..., :3] + 1e-6                    # PairwiseDistance eps
    dist = diff.pow(2).sum(-1, keepdim=True).sqrt()
    wn = dist / dist.sum(dim=2, keepdim=True)
    return torch.cat([grouped[..., :3] - kp, keyfeats * wn], dim=3)


# --------------------------------------------------------------------------- #
# candidates, KNN, target-side concat
# --------------------------------------------------------------------------- #
def voxelize(centres, r, s, G=None):
    """voxelize.py:19-83, scalar float64 formula."""
    return native.candidates(centres, r, s, G)


def knn(ref, query, k):
    """knn_cuda.KNN(k, transpose_mode=True) contract."""
    return native.knn(ref, query, k)


def cat_feat_tgt(cand, tgt_xyz, tgt_feat, dist, idx, per_feature_weight=True):
    """get_cat_feat_tgt.py:53-96 incl. the per-FEATURE weight (Q7;
    per_feature_weight=False: the per-neighbour weight of the commented line :64).
    cand [B,M,C,3] f32, dist/idx [B,M*C,K]. float64 result [B,M,C,K,3+F]."""
    B, M, C, _ = cand.shape
    K = idx.shape[-1]
    Fd = tgt_feat.shape[-1]
    assert K == Fd, "the reference's view needs k_nn == C_deep_feat"
    w = dist / dist.sum(dim=2, keepdim=True, dtype=torch.float64)      # f64 [B,Q,K]
    feat = index_points(tgt_feat, idx).view(B, M, C, K, Fd)
    xyz = index_points(tgt_xyz, idx).view(B, M, C, K, 3)
    local = xyz - cand.unsqueeze(3)
    wmap = w.view(B, M, C, 1, K) if per_feature_weight else w.view(B, M, C, K, 1)
    return torch.cat([local, feat * wmap], dim=4)        # f32 local part promoted to f64


# --------------------------------------------------------------------------- #
# DFE and CPG
# --------------------------------------------------------------------------- #
def feat_embedding(sd, X):
    """deep_feat_embedding.py:29-60: three affine maps, max over
    the neighbour axis (-2)."""
    X = X.float()
    X = F.linear(X, sd["DFE.fc1.weight"], sd["DFE.fc1.bias"])
    X = F.linear(X, sd["DFE.fc2.weight"], sd["DFE.fc2.bias"])
    X = F.linear(X, sd["DFE.fc3.weight"], sd["DFE.fc3.bias"])
    return X.max(dim=-2)[0]


def cpg(sd, src_dfe, tgt_dfe_cf, cand, G, reshape_quirk=True):
    """cpg.py:27-60.

    src_dfe [B,M,32]; tgt_dfe_cf [B,M,C,32] as the DFE produces it (candidate,
    feature). The reference permutes it to [B,M,32,C] (deepVCP.py:106) and then
    re-reads that logical order as (C,32) (cpg.py:34): T'[c',f'] =
    tgt[(c'*32+f') % C, (c'*32+f') // C]  (SURVEY Q4). reshape_quirk=False: the cost
    volume the code intends, cost[c, f] = (src[f] - tgt[c, f])^2.
    """
    B, M, C, Fd = tgt_dfe_cf.shape
    assert C == G * G * G
    scr = tgt_dfe_cf.permute(0, 1, 3, 2).reshape(B, M, C, Fd) if reshape_quirk else tgt_dfe_cf
    cost = (src_dfe.view(B, M, 1, Fd) - scr).square()
    x = cost.view(B * M, G, G, G, Fd).permute(0, 4, 1, 2, 3)
    x = F.conv3d(x, sd["cpg.conv1.weight"], sd["cpg.conv1.bias"], padding=1)
    x = F.conv3d(x, sd["cpg.conv2.weight"], sd["cpg.conv2.bias"], padding=1)
    x = F.conv3d(x, sd["cpg.conv3.weight"], sd["cpg.conv3.bias"], padding=1)
    logits = x.reshape(B, M, C)
    w = torch.softmax(logits, dim=-1).unsqueeze(-1)
    vcp = (w * cand).sum(-2) / w.expand(-1, -1, -1, 3).sum(-2)
    return vcp, logits


# --------------------------------------------------------------------------- #
# pose solve (deepVCP_loss.py:13-90)
# --------------------------------------------------------------------------- #
def get_rigid_transform(x, y, reflection_fix=False, weights=None):
    """deepVCP_loss.py:13-44: R = V U^T, no reflection fix (Q10).
    Beyond the reference: reflection_fix=True applies the Z = diag(1, 1, det)
    the reference builds but never uses (:36-40); weights [B,n] gives the weighted solve."""
    if weights is None:
        cx = x.mean(dim=2, keepdim=True)
        cy = y.mean(dim=2, keepdim=True)
        H = (x - cx) @ (y - cy).transpose(1, 2)
    else:
        w = weights.to(x.dtype).unsqueeze(1)               # [B,1,n]
        cx = (x * w).sum(dim=2, keepdim=True) / w.sum(dim=2, keepdim=True)
        cy = (y * w).sum(dim=2, keepdim=True) / w.sum(dim=2, keepdim=True)
        H = ((x - cx) * w) @ (y - cy).transpose(1, 2)
    U, S, Vh = torch.linalg.svd(H)
    V = Vh.transpose(1, 2)
    R = V @ U.transpose(1, 2)
    if reflection_fix:
        Z = torch.diag_embed(torch.stack([torch.ones_like(S[:, 0]), torch.ones_like(S[:, 0]),
                                          torch.sign(torch.linalg.det(R))], dim=1))
        R = V @ Z @ U.transpose(1, 2)
    t = cy - R @ cx
    return R, t


def svd_optimization(x, y_pred, R_true, t_true, inlier_ratio=0.8, reflection_fix=False):
    """deepVCP_loss.py:57-90. x, y_pred [B,3,n] float64."""
    y_true = R_true @ x + t_true
    y_pred = y_pred.double()
    n = y_pred.shape[2]
    R1, t1 = get_rigid_transform(x, y_pred, reflection_fix)
    y1 = R1 @ x + t1
    d, _ = native.knn(y1.transpose(1, 2).float().contiguous(),
                      y_true.transpose(1, 2).float().contiguous(), 1)   # [B,n,1]
    d = d.squeeze(-1)
    keep = int(n * inlier_ratio)
    B = x.shape[0]
    inl = torch.empty(B, keep, dtype=torch.int64)
    for b in range(B):
        order = sorted(range(n), key=lambda i: (float(d[b, i]), i))
        inl[b] = torch.tensor(order[:keep])
    gi = inl.unsqueeze(1).expand(-1, 3, -1)
    y1i = torch.gather(y1, 2, gi)
    x1 = torch.gather(x, 2, gi)
    R2, t2 = get_rigid_transform(x1, y1i, reflection_fix)
    return R2, t2, R1, t1, inl


# --------------------------------------------------------------------------- #
# whole forward (deepVCP.py:24-110)
# --------------------------------------------------------------------------- #
def grid_size(r, s):
    return int(math.ceil((2 * r + s / 2) / s - 1e-9))


QUIRK_KEYPOINT_VIEW, QUIRK_PER_FEATURE_WEIGHT, QUIRK_COST_VOLUME_RESHAPE, QUIRK_IGNORE_T_INIT, \
    QUIRK_NO_REFLECTION_FIX, QUIRK_FPS_ORDER_MISMATCH = 1, 2, 4, 8, 16, 32
QUIRKS_REFERENCE = 63


def deepvcp_forward(sd, src_pts, tgt_pts, R_init, r, s, starts, k_topk=64, nsample=32,
                    fe_radius=0.1, fe_nsample=256, topk_override=None, quirks=QUIRKS_REFERENCE, t_init=None,
                    chained_fe=False):
    """One call = B independent B=1 forwards.

    quirks (default: all set = the reference as it runs): a clear bit selects the semantics the
    reference's code intends -- nothing of the reference pins those.

    src_pts/tgt_pts [B,C_in,N]; R_init [B,3,3] float64; starts = (src_start[B],
    kp_start[B], tgt_start[B]) the three FPS start draws in the reference's order
    (deepVCP.py:29,54,72). Returns a dict of every stage boundary.
    """
    B, C_in, N = src_pts.shape
    o = {}
    o["src_fe_xyz"], o["src_fe_feat"], o["src_fps"] = feat_extraction(
        sd, src_pts, starts[0], radius=fe_radius, nsample=fe_nsample, chained=chained_fe)
    o["scores"] = weighting_scores(sd, o["src_fe_feat"])
    o["topk_idx"] = topk_indices(o["scores"], k_topk) if topk_override is None else topk_override
    # Q5 (deepVCP.py:35,46,61; get_cat_feat_tgt.py:85): the feature tables are in FPS order, the reference addresses them
    # with original-order / key-point-local indices. Clear bit: key-point k is point fps[topk[k]], its neighbourhood
    # features are the key-points' own rows, the target features are addressed by original point index.
    mismatch = bool(quirks & QUIRK_FPS_ORDER_MISMATCH)
    rows = o["topk_idx"] if mismatch else torch.gather(o["src_fps"], 1, o["topk_idx"])
    kp = gather_keypoints(src_pts, rows, bool(quirks & QUIRK_KEYPOINT_VIEW))   # [B,64,C_in]
    o["src_keypts_full"] = kp
    kp_xyz = kp[:, :, :3].contiguous()
    new_xyz, grouped, picked, kp_fps = sample_and_group(k_topk, 1, nsample, kp_xyz, None, starts[1])
    o["picked_idx"], o["kp_fps"], o["src_grouped"] = picked, kp_fps, grouped
    keyfeats = index_points(o["src_fe_feat"] if mismatch else index_points(o["src_fe_feat"], o["topk_idx"]), picked)
    o["src_cat"] = cat_feat_src(kp, grouped, keyfeats)
    tgt_xyz = tgt_pts[:, :3, :].permute(0, 2, 1).contiguous()
    o["tgt_fe_xyz"], o["tgt_fe_feat"], o["tgt_fps"] = feat_extraction(
        sd, tgt_pts, starts[2], radius=fe_radius, nsample=fe_nsample, chained=chained_fe)
    centres = (R_init @ kp_xyz.transpose(1, 2).double()).transpose(1, 2).contiguous()  # Q6: no t
    if t_init is not None and not (quirks & QUIRK_IGNORE_T_INIT):
        centres = centres + t_init.double().reshape(-1, 1, 3)
    o["centres"] = centres
    G = grid_size(r, s)
    cand = voxelize(centres, r, s, G)
    o["candidates"] = cand
    Q = cand.shape[1] * cand.shape[2]
    dist, idx = knn(tgt_xyz, cand.reshape(B, Q, 3), nsample)
    o["knn_dist"], o["knn_idx"] = dist, idx
    tfeat = o["tgt_fe_feat"]
    if not mismatch:
        tfeat = torch.zeros_like(tfeat).scatter(1, o["tgt_fps"].unsqueeze(-1).expand(-1, -1, tfeat.shape[-1]), tfeat)
    tgt_cat = cat_feat_tgt(cand, tgt_xyz, tfeat, dist, idx, bool(quirks & QUIRK_PER_FEATURE_WEIGHT))
    o["src_dfe"] = feat_embedding(sd, o["src_cat"])                   # [B,64,32]
    o["tgt_dfe"] = feat_embedding(sd, tgt_cat)                        # [B,64,C,32]
    del tgt_cat
    o["vcp"], o["logits"] = cpg(sd, o["src_dfe"], o["tgt_dfe"], cand, G, bool(quirks & QUIRK_COST_VOLUME_RESHAPE))
    o["src_keypts"] = kp_xyz
    return o


def pose_from_forward(src_keypts, vcp, R_true, t_true, quirks=QUIRKS_REFERENCE):
    """train.py:110 -> deepVCP_loss.py:105-121 (pose only)."""
    x = src_keypts.permute(0, 2, 1).double()
    y = vcp.permute(0, 2, 1).double()
    return svd_optimization(x, y, R_true.double(), t_true.double(),
                            reflection_fix=not (quirks & QUIRK_NO_REFLECTION_FIX))
