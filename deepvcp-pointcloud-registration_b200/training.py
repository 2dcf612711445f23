"""Training-mode forward and loss -- the step after the hot path.

Reference: train.py:93-138 (model.train(); forward; deepVCP_loss; backward; Adam),
deepVCP.py:24-110, deepVCP_loss.py:13-121.

What the reference differentiates and what it does not decides the split:

* The index-producing stages carry NO gradient in the reference either (integer outputs, or
  `torch.no_grad()` inside knn_cuda): farthest point sampling, ball query, the key-point
  grouping, the candidate lattice and both KNN searches. They run on the sm_100a kernels of
  this package (the stages that make the reference's training step slow: a Python loop of
  N rounds, an int64 sort of [B,S,N], an N x Q distance matrix).
* The differentiable stages are the reference's own tensor algebra written with torch
  operators on the device, so autograd returns the reference's gradients and BatchNorm in
  train mode uses and updates batch statistics exactly as nn.BatchNorm2d does
  (pointnet2_utils.py:196-198).
* One stage has a hand-written forward AND backward: the target-side embedding, whose input in the reference is
  a float64 [B,64,C,32,35] tensor (763 MB per pair at the KITTI shape, twice that with its autograd graph).
  `TargetEmbedding` evaluates it with the fused gather + embedding kernel and differentiates it with
  dvcp_dfe_tgt_backward (arg-max neighbours recomputed, gradients to the collapsed 32 x 35 map and to the target
  feature table); autograd carries the collapsed map's gradient on to fc1 / fc2 / fc3. `fused_embedding=False`
  selects the plain autograd form of the same stage.

`DeepVCP.forward` routes here when the module is in train mode; `deepVCP_loss` routes to
`loss()` when its prediction carries a gradient. Parameters that receive no gradient in the
reference receive none here (the weighting layer only produces indices, deepVCP.py:36).
"""
import contextlib

import torch
import torch.nn.functional as F

from . import functional as F_
from ._lib import cloud_pm
from ._lib import QUIRK_COST_VOLUME_RESHAPE, QUIRK_FPS_ORDER_MISMATCH, QUIRK_IGNORE_T_INIT, QUIRK_KEYPOINT_VIEW, QUIRK_PER_FEATURE_WEIGHT, require_cuda
from .get_cat_feat_src import Get_Cat_Feat_Src
from .knn_cuda import KNN
from .pointnet2_utils import farthest_point_sample, query_ball_point, sample_and_group
from .voxelize import voxelize


@contextlib.contextmanager
def fp32_math():
    """cuDNN / cuBLAS TF32 off inside the block: the reference's CPU gradients are FP32, and the convolutions of
    cpg.py would otherwise carry 1e-3-level noise (SURVEY 8c 'oracle numerics settings')."""
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def _gather_rows(points, idx):
    """points [B,N,C], idx [B,...] int64 -> [B,...,C]; differentiable in `points` (index_points, :43-60)."""
    B = points.shape[0]
    bi = torch.arange(B, device=points.device).view(B, *([1] * (idx.dim() - 1)))
    return points[bi, idx]


def set_abstraction(sa, xyz_cm, feats_cm, start=None):
    """PointNetSetAbstraction.forward in the module's current mode (pointnet2_utils.py:176-202):
    FPS + ball query on the kernels, grouping / shared MLP / max with torch operators.
    xyz_cm [B,3,N], feats_cm [B,D,N] or None -> new_xyz [B,S,3], features [B,S,C_out], fps_idx [B,S]."""
    xyz = xyz_cm.permute(0, 2, 1).contiguous()
    B, N, _ = xyz.shape
    S = sa.npoint
    with torch.no_grad():
        fps_idx = farthest_point_sample(xyz, S, start=start)
        new_xyz = _gather_rows(xyz, fps_idx)
        idx = query_ball_point(sa.radius, sa.nsample, xyz, new_xyz)
        if int(idx.max()) >= N:
            raise IndexError("index %d is out of bounds for dimension 1 with size %d" % (int(idx.max()), N))
        new_points = _gather_rows(xyz, idx) - new_xyz.view(B, S, 1, 3)
        if feats_cm is not None:
            new_points = torch.cat([new_points, _gather_rows(feats_cm.permute(0, 2, 1), idx)], dim=-1)
    x = new_points.permute(0, 3, 2, 1)                       # [B, 3+D, nsample, S]
    for conv, bn in zip(sa.mlp_convs, sa.mlp_bns):
        x = F.relu(bn(conv(x.float())))
    return new_xyz, torch.max(x, 2)[0].permute(0, 2, 1), fps_idx


def embedding(dfe, X):
    """feat_embedding_layer.forward (deep_feat_embedding.py:29-60): three Linear layers, max over the neighbour
    axis (dim -2) -- the same function for the source [B,K,ns,35] and target [B,K,C,ns,35] tensors."""
    X = dfe.fc3(dfe.fc2(dfe.fc1(X.float())))
    return torch.max(X, dim=-2)[0]


class TargetEmbedding(torch.autograd.Function):
    """out[b,q,c] = max_j (Wc x_j + bc)[c] over the 32 neighbours of candidate q (get_cat_feat_tgt.py:53-96 +
    deep_feat_embedding.py:46-60) as ONE kernel forward and ONE kernel backward. Differentiable inputs: the
    collapsed map Wc [32,35], its bias bc [32] (functions of fc1 / fc2 / fc3 on the autograd tape) and the target
    feature table [B,N,32]; everything else is data (knn_cuda runs under no_grad, :45-52)."""

    @staticmethod
    def forward(ctx, Wc, bc, tfeat, cand, tgt_xyz, dist, idx32, dfe_mod, quirks):
        B, N, _ = tfeat.shape
        ctx.save_for_backward(Wc, tfeat, cand, tgt_xyz, dist, idx32)
        ctx.dfe_mod, ctx.quirks = dfe_mod, quirks
        return F_.dfe_tgt_fused(cand, cloud_pm(tgt_xyz), tfeat, dist, idx32, B, N, dfe_mod.params(), quirks)

    @staticmethod
    def backward(ctx, gout):
        Wc, tfeat, cand, tgt_xyz, dist, idx32 = ctx.saved_tensors
        B, N, _ = tfeat.shape
        gw, gb, gf = F_.dfe_tgt_backward(cand, cloud_pm(tgt_xyz), tfeat, dist, idx32, B, N, ctx.dfe_mod.params(),
                                         Wc.detach(), ctx.quirks, gout.contiguous())
        return gw, gb, gf, None, None, None, None, None, None


def collapsed_embedding(dfe):
    """Wc = W3 W2 W1, bc = W3 (W2 b1 + b2) + b3 on the autograd tape (no activation between the layers)."""
    W1, W2, W3 = dfe.fc1.weight, dfe.fc2.weight, dfe.fc3.weight
    return W3 @ W2 @ W1, W3 @ (W2 @ dfe.fc1.bias + dfe.fc2.bias) + dfe.fc3.bias


def corresponding_points(cpg_mod, src_dfe, tgt_dfe, candidates, G, reshape_quirk):
    """cpg.forward (cpg.py:27-60). src_dfe [B,K,32], tgt_dfe [B,K,C,32], candidates [B,K,C,3] -> vcp [B,K,3]."""
    B, K, C, _ = candidates.shape
    if reshape_quirk:   # deepVCP.py:106 + cpg.py:34: the permuted tensor's logical [32, C] order re-read as [C, 32]
        tgt = tgt_dfe.permute(0, 1, 3, 2).reshape(B, K, G, G, G, 32)
    else:
        tgt = tgt_dfe.reshape(B, K, G, G, G, 32)
    cost = torch.square(src_dfe.reshape(B, K, 1, 1, 1, 32) - tgt)
    x = cost.permute(0, 1, 5, 2, 3, 4).flatten(0, 1)
    x = cpg_mod.conv3(cpg_mod.conv2(cpg_mod.conv1(x)))
    w = torch.softmax(x.reshape(B, K, C), dim=-1).unsqueeze(-1)
    return torch.sum(w * candidates, -2) / torch.sum(w.expand(-1, -1, -1, 3), -2)


def forward(model, src_pts, tgt_pts, R_init, t_init, starts=None, keep_stages=False, topk_override=None,
            fused_embedding=True):
    """DeepVCP.forward (deepVCP.py:24-110) with a gradient. Same arguments and results as the inference
    forward; B > 1 = B independent pairs (BatchNorm statistics are taken over the batch, as torch does)."""
    dev = model.cpg.conv1.weight.device
    if dev.type != "cuda":
        raise RuntimeError("DeepVCP (b200) needs its parameters on a CUDA device; there is no CPU fallback")
    if model.FE1.chained:
        raise NotImplementedError("training with chained_fe")
    src = src_pts.to(dev)
    tgt = tgt_pts.to(dev)
    R = R_init.to(dev)
    require_cuda(src, tgt, R)
    B, C_in, N = src.shape
    K, ns, q = model.K_topk, model.nsample, model.quirks
    if starts is None:
        starts = model.draw_starts(B, N)
    sa = model.FE1.sa1
    feats = lambda p: p[:, 3:, :] if C_in > 3 else None
    # feature extraction of the source (deepVCP.py:29; sa1 only: SURVEY Q1)
    _, sfeat, sfps = set_abstraction(sa, src[:, :3, :], feats(src), starts[0])
    # key-point choice (weighting_layer.py:26-33): indices only, no gradient
    with torch.no_grad():
        scores = model.WL.fc3(model.WL.fc2(model.WL.fc1(sfeat)))
        topk = (torch.topk(scores, K, dim=1).indices.view(B, K) if topk_override is None
                else topk_override.to(dev).view(B, K))
        intended_order = not (q & QUIRK_FPS_ORDER_MISMATCH)                             # Q5, see DeepVCP.match
        rows = torch.gather(sfps, 1, topk) if intended_order else topk
        g = torch.gather(src, 2, rows.view(B, 1, K).expand(-1, C_in, -1))               # [B,C_in,K]
        keypts = g.reshape(B, K, C_in) if q & QUIRK_KEYPOINT_VIEW else g.permute(0, 2, 1).contiguous()   # Q3
        # grouping among the key-points (deepVCP.py:54-56)
        _, grouped, picked = sample_and_group(K, model.group_radius, ns, keypts[:, :, :3].contiguous(), None,
                                              returnidx=True, start=starts[1])
    src_keyfeats = _gather_rows(_gather_rows(sfeat, topk) if intended_order else sfeat, picked)   # deepVCP.py:61 (Q5)
    src_cat = Get_Cat_Feat_Src()(keypts, grouped, src_keyfeats)
    # target side
    tgt_xyz = tgt[:, :3, :].permute(0, 2, 1).contiguous()
    _, tfeat, tfps = set_abstraction(sa, tgt[:, :3, :], feats(tgt), starts[2])
    with torch.no_grad():
        centres = torch.matmul(R.double(), keypts[:, :, :3].permute(0, 2, 1).double()).permute(0, 2, 1)
        if not (q & QUIRK_IGNORE_T_INIT) and t_init is not None:                       # Q6
            centres = centres + t_init.to(dev).double().reshape(1, 1, 3)
        cand = voxelize(centres.contiguous(), model.r, model.s)                         # [B,K,C,3] float32
        C = cand.shape[2]
        G = round(C ** (1.0 / 3.0))
        dist, idx = KNN(k=ns, transpose_mode=True)(tgt_xyz, cand.view(B, K * C, 3))     # get_cat_feat_tgt.py:45-52
    if intended_order:   # feature rows by original point index, like the KNN indices that address them
        tfeat_used = torch.zeros_like(tfeat).scatter(1, tfps.unsqueeze(-1).expand(-1, -1, 32), tfeat)
    else:
        tfeat_used = tfeat
    src_dfe = embedding(model.DFE, src_cat)
    if fused_embedding:
        # gather + weights + embedding + max in one kernel each way; the float64 [B,K,C,32,35] tensor is never built
        Wc, bc = collapsed_embedding(model.DFE)
        tgt_dfe = TargetEmbedding.apply(Wc, bc, tfeat_used, cand.view(B, K * C, 3), tgt_xyz.float().contiguous(), dist,
                                        idx.int(), model.DFE, q).view(B, K, C, 32)
    else:
        with torch.no_grad():
            w = dist / torch.sum(dist, dim=2, keepdim=True, dtype=torch.float64)        # :57-58, float64
            local = _gather_rows(tgt_xyz, idx).view(B, K, C, ns, 3) - cand.unsqueeze(3)  # :86-89
        picked_feat = _gather_rows(tfeat_used, idx).view(B, K, C, ns, 32)               # :85
        if q & QUIRK_PER_FEATURE_WEIGHT:                                                # :65,92 (Q7): weight by feature
            wmap = w.view(B, K, C, 1, ns)
        else:
            wmap = w.view(B, K, C, ns, 1)
        tgt_cat = torch.cat((local, picked_feat * wmap), dim=4)                         # float64, :93-96
        tgt_dfe = embedding(model.DFE, tgt_cat)
    vcp = corresponding_points(model.cpg, src_dfe, tgt_dfe, cand, G, bool(q & QUIRK_COST_VOLUME_RESHAPE))
    if keep_stages:
        model.last = dict(src_fps=sfps, tgt_fps=tfps, src_fe_feat=sfeat, tgt_fe_feat=tfeat, scores=scores, topk_idx=topk,
                          src_keypts_full=keypts, picked_idx=picked, src_cat=src_cat, src_dfe=src_dfe,
                          centres=centres, candidates=cand, knn_dist=dist, knn_idx=idx, tgt_dfe=tgt_dfe, vcp=vcp)
    return keypts[:, :, :3], vcp


def rigid_transform(x, y):
    """get_rigid_transform with a gradient (deepVCP_loss.py:13-44): torch.svd of the 3x3 covariance, R = V U^T,
    no reflection correction (Q10)."""
    cx, cy = x.mean(dim=2, keepdim=True), y.mean(dim=2, keepdim=True)
    H = torch.matmul(x - cx, (y - cy).permute(0, 2, 1))
    U, _, V = torch.svd(H)
    Rm = torch.matmul(V, U.permute(0, 2, 1))
    return Rm, cy + torch.matmul(-Rm, cx)


def loss(x, y_pred, R_true, t_true, alpha):
    """deepVCP_loss (deepVCP_loss.py:57-121) with a gradient. x, y_pred [B,n,3] -> (loss, R [B,3,3], t [B,3,1]).
    The 1-NN search of the outlier rejection (:70-72) runs on the KNN kernel; everything else is autograd."""
    B, n, _ = y_pred.shape
    dev = y_pred.device
    x = x.to(dev).permute(0, 2, 1).double()
    y = y_pred.permute(0, 2, 1).double()
    Rt = R_true.to(dev).double()
    tt = t_true.to(dev).double().reshape(B, 3, -1)
    y_true = torch.matmul(Rt, x) + tt
    R1, t1 = rigid_transform(x, y)
    y1 = torch.matmul(R1, x) + t1
    d, _ = KNN(k=1, transpose_mode=False)(y1.detach(), y_true)                          # [B,1,n]
    inl = torch.topk(d.to(dev), k=int(n * 0.8), dim=-1, largest=False, sorted=True).indices.expand(-1, 3, -1)
    y1i, x1 = torch.gather(y1, -1, inl), torch.gather(x, -1, inl)
    R2, t2 = rigid_transform(x1, y1i)
    y2 = torch.matmul(R2, x1) + t2
    y_true_i = torch.matmul(Rt, x1) + tt
    l1 = torch.mean(torch.abs(y_true_i - y2))
    l2 = torch.abs(torch.mean(y2 - y_true_i))
    return alpha * l1 + (1 - alpha) * l2, R2, t2
