"""Tensor-level wrappers of the C ABI: allocate outputs with torch, pass raw
pointers and the current stream, raise on any non-zero return code.

Clouds may be passed either point-major `[B, N, 3]` or channel-major
`[B, C, N]` views with arbitrary strides; they are read in place.
"""
import math

import numpy as np
import torch

from . import _lib
from ._lib import (Cloud, CloudIndex, MlpLayer, NULL_CLOUD, NULL_INDEX, check, cloud_cm, cloud_pm, lib, ptr,
                   require_cuda, stream_ptr)


# number of kernels of libdvcp_b200.so launched through this module (bench.py's
# `gpu_launches`); every wrapper adds what its entry point launches.
LAUNCHES = 0

# Spatially pruned kernels (exact; results identical to the brute-force kernels).
# Tests flip this to compare the two families.
USE_INDEX = True


def _count(n):
    global LAUNCHES
    LAUNCHES += n


def _f32c(t):
    if t.dtype != torch.float32:
        raise RuntimeError("expected a float32 tensor, got %s" % t.dtype)
    return t if t.is_contiguous() else t.contiguous()


def _starts_to_device(start, B, device):
    start = torch.as_tensor(start, dtype=torch.int64).reshape(-1)
    if start.numel() != B:
        raise RuntimeError("need one FPS start index per cloud")
    return start.to(device, non_blocking=True)


def draw_fps_start(B: int, N: int) -> torch.Tensor:
    """The reference draws the first FPS index on the CPU default generator
    (pointnet2_utils.py:75); doing the same keeps torch.manual_seed runs aligned."""
    return torch.randint(0, N, (B,), dtype=torch.long)


# --------------------------------------------------------------------------- #
class SpatialIndex:
    """Device buffers of a cloud's spatial index (dvcp_cloud_index_t). big=True admits clouds of up to
    131072 points (multi-CTA build; consumed by the KNN only)."""

    def __init__(self, B, N, device, big=False):
        self.cap = lib().dvcp_index_capacity_any(N) if big else lib().dvcp_index_capacity(N)
        if self.cap == 0:
            raise RuntimeError("clouds of %d points cannot be indexed (64..%d)" % (N, 131072 if big else 16384))
        self.sorted_pt = torch.empty(B, self.cap, 4, dtype=torch.float32, device=device)   # x, y, z, index bits
        self.bucket_box = torch.empty(B, self.cap // 32, 8, dtype=torch.float32, device=device)
        self.B = B

    def c(self, lo=0):
        """ctypes view starting at batch item `lo`."""
        return CloudIndex(self.sorted_pt[lo:].data_ptr(), self.bucket_box[lo:].data_ptr(), self.cap)

    @staticmethod
    def indexable(N, dtype=torch.float32):
        return USE_INDEX and dtype == torch.float32 and lib().dvcp_index_capacity(N) > 0

    @staticmethod
    def knn_indexable(N, dtype=torch.float32):
        return USE_INDEX and dtype == torch.float32 and lib().dvcp_index_capacity_any(N) > 0


def fps(xyz_cloud: Cloud, device, dtype, B, N, npoint, start, want64=True, want32=False, index=None):
    start = _starts_to_device(start, B, device)
    if index is None and N > 2048 and SpatialIndex.indexable(N, dtype):
        index = SpatialIndex(B, N, device)   # workspace: lets the library spread each cloud over a cluster
    o64 = torch.empty(B, npoint, dtype=torch.int64, device=device) if want64 else None
    o32 = torch.empty(B, npoint, dtype=torch.int32, device=device) if want32 else None
    code = lib().dvcp_fps(xyz_cloud, 0 if dtype == torch.float32 else 1, B, N, npoint, ptr(start), ptr(o64),
                          ptr(o32), index.c() if index is not None else NULL_INDEX, stream_ptr(device))
    check(code, "dvcp_fps")
    _count(2 if index is not None and N > 2048 else 1)   # index build + cluster kernel (an upper bound for huge B)
    return o64, o32


def fps_indexed(xyz_cloud: Cloud, device, B, N, npoint, start, index, want64=False, want32=True, concurrent=False):
    """FPS of float32 clouds whose spatial index is already built (the index is only read)."""
    start = _starts_to_device(start, B, device)
    o64 = torch.empty(B, npoint, dtype=torch.int64, device=device) if want64 else None
    o32 = torch.empty(B, npoint, dtype=torch.int32, device=device) if want32 else None
    check(lib().dvcp_fps_indexed(xyz_cloud, B, N, npoint, ptr(start), ptr(o64), ptr(o32), index.c(),
                                 int(concurrent), stream_ptr(device)), "dvcp_fps_indexed")
    _count(1)
    return o64, o32


def gather_rows(points, idx32):
    """points [B,N,C] float32 (C % 4 == 0), idx32 [B,M] int32 -> [B,M,C]."""
    require_cuda(points, idx32)
    B, N, C = points.shape
    M = idx32.shape[1]
    out = torch.empty(B, M, C, dtype=torch.float32, device=points.device)
    check(lib().dvcp_index_points_i32(ptr(_f32c(points)), ptr(idx32.contiguous()), B, N, C, M, ptr(out),
                                      stream_ptr(points.device)), "dvcp_index_points_i32")
    _count(1)
    return out


def build_index(xyz_cloud: Cloud, device, B, N, big=False):
    """Spatial index of B float32 clouds. big=True: any 64 <= N <= 131072 (above 16384 points the
    index is built by several CTAs per cloud and serves the KNN only)."""
    index = SpatialIndex(B, N, device, big=big)
    if not big or index.cap <= 16384:
        check(lib().dvcp_build_index(xyz_cloud, B, N, index.c(), stream_ptr(device)), "dvcp_build_index")
        _count(1)
        return index
    nbytes = lib().dvcp_build_index_workspace_bytes(B, N)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
    check(lib().dvcp_build_index_ws(xyz_cloud, B, N, index.c(), ptr(ws), nbytes, stream_ptr(device)),
          "dvcp_build_index_ws")
    _count(3 + int(math.log2(index.cap // 16384)))   # bbox, run sort, merges, publish
    return index


def fps_plain(xyz_pm, npoint, start):
    require_cuda(xyz_pm)
    B, N, _ = xyz_pm.shape
    start = _starts_to_device(start, B, xyz_pm.device)
    out = torch.empty(B, npoint, dtype=torch.int64, device=xyz_pm.device)
    check(lib().dvcp_fps_plain(cloud_pm(xyz_pm), B, N, npoint, ptr(start), ptr(out), stream_ptr(xyz_pm.device)),
          "dvcp_fps_plain")
    _count(1)
    return out


def square_distance(src_pm, dst_pm):
    """float32 clouds -> float32; a float64 cloud on either side -> double (torch's promotion)."""
    require_cuda(src_pm, dst_pm)
    B, S, _ = src_pm.shape
    N = dst_pm.shape[1]
    if src_pm.dtype == torch.float64 or dst_pm.dtype == torch.float64:
        src_pm, dst_pm = src_pm.double(), dst_pm.double()
        out = torch.empty(B, S, N, dtype=torch.float64, device=src_pm.device)
        check(lib().dvcp_square_distance_f64(cloud_pm(src_pm), cloud_pm(dst_pm), B, S, N, ptr(out),
                                             stream_ptr(src_pm.device)), "dvcp_square_distance_f64")
        _count(1)
        return out
    out = torch.empty(B, S, N, dtype=torch.float32, device=src_pm.device)
    check(lib().dvcp_square_distance(cloud_pm(src_pm), cloud_pm(dst_pm), B, S, N, ptr(out),
                                     stream_ptr(src_pm.device)), "dvcp_square_distance")
    _count(1)
    return out


def radius2_f32(radius) -> float:
    """float32(radius ** 2): the comparison constant of pointnet2_utils.py:102."""
    return float(np.float32(float(radius) ** 2))


def ball_query(radius, nsample, xyz_pm, new_xyz_pm):
    require_cuda(xyz_pm, new_xyz_pm)
    B, N, _ = xyz_pm.shape
    S = new_xyz_pm.shape[1]
    out = torch.empty(B, S, nsample, dtype=torch.int64, device=xyz_pm.device)
    if xyz_pm.dtype == torch.float64 or new_xyz_pm.dtype == torch.float64:
        # double arithmetic, bound = the Python double radius**2 (pointnet2_utils.py:102)
        xyz_pm, new_xyz_pm = xyz_pm.double(), new_xyz_pm.double()
        check(lib().dvcp_ball_query_f64(cloud_pm(xyz_pm), cloud_pm(new_xyz_pm), B, N, S, float(radius) ** 2, nsample,
                                        ptr(out), stream_ptr(xyz_pm.device)), "dvcp_ball_query_f64")
        _count(1)
        return out
    check(lib().dvcp_ball_query(cloud_pm(xyz_pm), cloud_pm(new_xyz_pm), B, N, S, radius2_f32(radius), nsample,
                                ptr(out), stream_ptr(xyz_pm.device)), "dvcp_ball_query")
    _count(1)
    return out


def index_points(points, idx):
    require_cuda(points, idx)
    B, N, C = points.shape
    points = _f32c(points)
    idx = idx.contiguous().to(torch.int64)
    M = idx[0].numel()
    out = torch.empty(*idx.shape, C, dtype=torch.float32, device=points.device)
    check(lib().dvcp_index_points(ptr(points), ptr(idx), B, N, C, M, ptr(out), stream_ptr(points.device)),
          "dvcp_index_points")
    _count(1)
    return out


class FoldedMlp:
    """Device copies of the folded shared-MLP parameters (conv1x1 + eval BN)."""

    def __init__(self, convs, bns):
        self.tensors = []
        self.layers = (MlpLayer * len(convs))()
        for i, (conv, bn) in enumerate(zip(convs, bns)):
            W = conv.weight.detach().reshape(conv.out_channels, conv.in_channels).float().contiguous()
            b = conv.bias.detach().float().contiguous()
            invstd = torch.rsqrt(bn.running_var.detach().float() + bn.eps)
            alpha = (bn.weight.detach().float() * invstd).contiguous()
            beta = (bn.bias.detach().float() - bn.running_mean.detach().float() * alpha).contiguous()
            self.tensors += [W, b, alpha, beta]
            self.layers[i] = MlpLayer(W.data_ptr(), b.data_ptr(), alpha.data_ptr(), beta.data_ptr(),
                                      conv.in_channels, conv.out_channels)
        self.n = len(convs)
        self.out_ch = convs[-1].out_channels


def sa_layer(xyz_cloud, feats_cloud, D, centroid_idx32, B, N, S, radius, nsample, mlp: FoldedMlp, device,
             want_xyz=True, index=None):
    out = torch.empty(B, S, mlp.out_ch, dtype=torch.float32, device=device)
    oxyz = torch.empty(B, S, 3, dtype=torch.float32, device=device) if want_xyz else None
    ws = torch.empty(B * S, dtype=torch.uint8, device=device) if index is not None else None
    code = lib().dvcp_sa_layer(xyz_cloud, feats_cloud if D > 0 else NULL_CLOUD, D, ptr(centroid_idx32), B, N, S,
                               radius2_f32(radius), nsample, mlp.layers, mlp.n,
                               index.c() if index is not None else NULL_INDEX, ptr(ws), ptr(out), ptr(oxyz),
                               stream_ptr(device))
    check(code, "dvcp_sa_layer")
    _count(2 if index is not None else 1)
    return oxyz, out


def sa_layer_f64(xyz_cloud, feats_cloud, D, centroid_idx32, B, N, S, radius, nsample, mlp: FoldedMlp, device,
                 want_xyz=True):
    """SA layer of float64 clouds (xyz and features clouds of doubles): features float32 [B,S,C], xyz float64."""
    out = torch.empty(B, S, mlp.out_ch, dtype=torch.float32, device=device)
    oxyz = torch.empty(B, S, 3, dtype=torch.float64, device=device) if want_xyz else None
    check(lib().dvcp_sa_layer_f64(xyz_cloud, feats_cloud if D > 0 else NULL_CLOUD, D, ptr(centroid_idx32), B, N, S,
                                  float(radius) ** 2, nsample, mlp.layers, mlp.n, ptr(out), ptr(oxyz),
                                  stream_ptr(device)), "dvcp_sa_layer_f64")
    _count(1)
    return oxyz, out


def sa_layer_all(xyz_cloud, feats_cloud, D, identity_idx32, B, N, radius, nsample, mlp: FoldedMlp, device, index):
    """SA features of every point in original order: [B,N,out_ch]."""
    out = torch.empty(B, N, mlp.out_ch, dtype=torch.float32, device=device)
    ws = torch.empty(B * N + 32, dtype=torch.uint8, device=device)
    code = lib().dvcp_sa_layer_all(xyz_cloud, feats_cloud if D > 0 else NULL_CLOUD, D, ptr(identity_idx32), B, N,
                                   radius2_f32(radius), nsample, mlp.layers, mlp.n, index.c(), ptr(ws), ptr(out),
                                   stream_ptr(device))
    check(code, "dvcp_sa_layer_all")
    _count(2)
    return out


def linear_rows(X, W, b):
    """X [..., in] float32 -> [..., out] = X W^T + b (out <= 32, in <= 128)."""
    require_cuda(X, W, b)
    X = _f32c(X)
    out_ch, in_ch = W.shape
    rows = X.numel() // in_ch
    Y = torch.empty(*X.shape[:-1], out_ch, dtype=torch.float32, device=X.device)
    check(lib().dvcp_linear_rows(ptr(X), rows, in_ch, out_ch, ptr(_f32c(W.detach())), ptr(_f32c(b.detach())), ptr(Y),
                                 stream_ptr(X.device)), "dvcp_linear_rows")
    _count(1)
    return Y


def weighting_scores(X, W1, b1, W2, b2, W3, b3):
    require_cuda(X)
    B, S, _ = X.shape
    X = _f32c(X)
    scores = torch.empty(B, S, dtype=torch.float32, device=X.device)
    check(lib().dvcp_weighting_scores(ptr(X), B, S, ptr(W1), ptr(b1), ptr(W2), ptr(b2), ptr(W3), ptr(b3),
                                      ptr(scores), stream_ptr(X.device)), "dvcp_weighting_scores")
    _count(1)
    return scores


def topk(scores, K):
    require_cuda(scores)
    B, S = scores.shape
    scores = _f32c(scores)
    out = torch.empty(B, K, dtype=torch.int64, device=scores.device)
    check(lib().dvcp_topk(ptr(scores), B, S, K, ptr(out), stream_ptr(scores.device)), "dvcp_topk")
    _count(1)
    return out


def keypoint_stage(src_pts, topk_idx, kp_start, src_feat, R_init, radius, nsample, dfe, quirks,
                   want_cat=False, want_picked=False, t_init=None):
    """src_pts [B,C_in,N] float32 or float64 (the key-points come back in that dtype, like the reference's)."""
    require_cuda(src_pts, topk_idx, src_feat, R_init)
    B, C_in, N = src_pts.shape
    Kp = topk_idx.shape[1]
    S = src_feat.shape[1]
    dev = src_pts.device
    f64 = src_pts.dtype == torch.float64
    src_pts = src_pts.contiguous() if f64 else _f32c(src_pts)
    src_feat = _f32c(src_feat)
    R_init = R_init.to(torch.float64).contiguous()
    kp_start = _starts_to_device(kp_start, B, dev)
    t_dev, t_stride = None, 0
    if t_init is not None and not (quirks & _lib.QUIRK_IGNORE_T_INIT):
        t_dev = torch.as_tensor(t_init).to(dev, torch.float64).reshape(-1, 3).contiguous()
        if t_dev.shape[0] not in (1, B):
            raise RuntimeError("t_init must be [1,3] or [B,3]")
        t_stride = 3 if t_dev.shape[0] == B and B > 1 else 0
    keypts = torch.empty(B, Kp, C_in, dtype=src_pts.dtype, device=dev)
    picked = torch.empty(B, Kp, nsample, dtype=torch.int64, device=dev) if want_picked else None
    cat = torch.empty(B, Kp, nsample, 35, dtype=torch.float32, device=dev) if want_cat else None
    sdfe = torch.empty(B, Kp, 32, dtype=torch.float32, device=dev)
    centres = torch.empty(B, Kp, 3, dtype=torch.float64, device=dev)
    fn = lib().dvcp_keypoint_stage_f64 if f64 else lib().dvcp_keypoint_stage
    r2 = float(radius) ** 2 if f64 else radius2_f32(radius)
    code = fn(ptr(src_pts), C_in, B, N, ptr(topk_idx.contiguous()), Kp, ptr(kp_start),
              ptr(src_feat), S, ptr(R_init), ptr(t_dev), t_stride, r2, nsample, dfe, quirks,
              ptr(keypts), ptr(picked), ptr(cat), ptr(sdfe), ptr(centres), stream_ptr(dev))
    check(code, "dvcp_keypoint_stage")
    _count(1)
    return keypts, picked, cat, sdfe, centres


def grid_size(r, s) -> int:
    g = lib().dvcp_grid_size(float(r), float(s))
    if g <= 0:
        check(g, "dvcp_grid_size")
    return g


def candidates(centres, r, s, G=None):
    require_cuda(centres)
    centres = centres.to(torch.float64).contiguous()
    lead = centres.shape[:-1]
    M = centres.numel() // 3
    G = grid_size(r, s) if G is None else G
    out = torch.empty(*lead, G * G * G, 3, dtype=torch.float32, device=centres.device)
    check(lib().dvcp_candidates(ptr(centres), M, float(r), float(s), G, ptr(out), stream_ptr(centres.device)),
          "dvcp_candidates")
    _count(1)
    return out


def knn(ref_cloud, device, B, N, query, K, want64=True, want32=False):
    require_cuda(query)
    query = _f32c(query)
    Q = query.shape[1]
    dist = torch.empty(B, Q, K, dtype=torch.float32, device=device)
    i64 = torch.empty(B, Q, K, dtype=torch.int64, device=device) if want64 else None
    i32 = torch.empty(B, Q, K, dtype=torch.int32, device=device) if want32 else None
    check(lib().dvcp_knn(ref_cloud, ptr(query), B, N, Q, K, ptr(dist), ptr(i64), ptr(i32), stream_ptr(device)),
          "dvcp_knn")
    _count(1)
    return dist, i64, i32


def knn_indexed(index, lo, device, B, N, query, K, chain=1, zline=0, want64=True, want32=False):
    """index: SpatialIndex whose batch items lo..lo+B-1 are the reference clouds."""
    require_cuda(query)
    query = _f32c(query)
    Q = query.shape[1]
    dist = torch.empty(B, Q, K, dtype=torch.float32, device=device)
    i64 = torch.empty(B, Q, K, dtype=torch.int64, device=device) if want64 else None
    i32 = torch.empty(B, Q, K, dtype=torch.int32, device=device) if want32 else None
    check(lib().dvcp_knn_indexed(index.c(lo), ptr(query), B, N, Q, K, chain, zline or chain, ptr(dist), ptr(i64),
                                 ptr(i32), stream_ptr(device)), "dvcp_knn_indexed")
    _count(1)
    return dist, i64, i32


def knn_groups(index, lo, device, B, N, query, K, group, zline, cell, pool_cap=0, want64=True, want32=False,
               stats=None):
    """KNN of query groups (one candidate lattice per key-point) through per-group shared-memory pools;
    index: SpatialIndex whose batch items lo..lo+B-1 are the reference clouds. Exact for any queries."""
    require_cuda(query)
    query = _f32c(query)
    Q = query.shape[1]
    dist = torch.empty(B, Q, K, dtype=torch.float32, device=device)
    i64 = torch.empty(B, Q, K, dtype=torch.int64, device=device) if want64 else None
    i32 = torch.empty(B, Q, K, dtype=torch.int32, device=device) if want32 else None
    nbytes = lib().dvcp_knn_groups_workspace_bytes(B, Q, group)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
    check(lib().dvcp_knn_groups(index.c(lo), ptr(query), B, N, Q, K, group, zline, float(cell), pool_cap, ptr(dist),
                                ptr(i64), ptr(i32), ptr(stats), ptr(ws), nbytes, stream_ptr(device)), "dvcp_knn_groups")
    _count(3)   # pool build + query kernel + the deferred index searches
    return dist, i64, i32


def dfe_tgt_fused(cand, tgt_cloud, tgt_feat, knn_dist, knn_idx32, B, N, dfe, quirks):
    require_cuda(cand, tgt_feat, knn_dist, knn_idx32)
    Q = knn_dist.shape[1]
    out = torch.empty(B, Q, 32, dtype=torch.float32, device=cand.device)
    check(lib().dvcp_dfe_tgt_fused(ptr(_f32c(cand)), tgt_cloud, ptr(_f32c(tgt_feat)), ptr(knn_dist), ptr(knn_idx32),
                                   B, N, Q, dfe, quirks, ptr(out), stream_ptr(cand.device)), "dvcp_dfe_tgt_fused")
    _count(1)
    return out


def dfe_tgt_backward(cand, tgt_cloud, tgt_feat, knn_dist, knn_idx32, B, N, dfe, w_collapsed, quirks, grad_out):
    """Gradient of dfe_tgt_fused (training): -> (grad of the collapsed map [32,35], grad of its bias [32],
    grad of tgt_feat [B,N,32])."""
    require_cuda(cand, tgt_feat, knn_dist, knn_idx32, grad_out)
    dev = cand.device
    Q = knn_dist.shape[1]
    gw = torch.zeros(32, 35, dtype=torch.float32, device=dev)
    gb = torch.zeros(32, dtype=torch.float32, device=dev)
    gf = torch.zeros(B, N, 32, dtype=torch.float32, device=dev)
    check(lib().dvcp_dfe_tgt_backward(ptr(_f32c(cand)), tgt_cloud, ptr(_f32c(tgt_feat)), ptr(knn_dist), ptr(knn_idx32), B, N, Q,
                                      dfe, ptr(_f32c(w_collapsed)), quirks, ptr(_f32c(grad_out)), ptr(gw), ptr(gb), ptr(gf),
                                      stream_ptr(dev)), "dvcp_dfe_tgt_backward")
    _count(1)
    return gw, gb, gf


def dfe_tc_operand(W1, b1, W2, b2, W3, b3, device):
    """Collapse the three un-activated Linear layers (float64) and lay the 32 x 40
    operand out for the tensor-core kernel; returns (b_hi, b_lo) device tensors."""
    W1, b1, W2, b2, W3, b3 = (t.detach().double().cpu() for t in (W1, b1, W2, b2, W3, b3))
    Wc = W3 @ W2 @ W1                                   # [32, 35]
    bc = W3 @ (W2 @ b1 + b2) + b3                       # [32]
    Bm = torch.zeros(32, 40, dtype=torch.float64)
    Bm[:, :32] = Wc[:, 3:35]
    Bm[:, 32:35] = Wc[:, 0:3]
    Bm[:, 35] = bc
    Bm = Bm.float()
    hi = (Bm.view(torch.int32) & -8192).view(torch.float32)       # clear the 13 low mantissa bits: exact TF32
    lo = Bm - hi
    n = torch.arange(32).view(32, 1).expand(32, 40)
    k = torch.arange(40).view(1, 40).expand(32, 40)
    off = ((n >> 3) * 10 + (k >> 2)) * 32 + (n & 7) * 4 + (k & 3)
    assert int(off[9, 17]) == lib().dvcp_dfe_tc_b_offset(9, 17) and lib().dvcp_dfe_tc_b_floats() == 1280
    img_hi = torch.zeros(1280, dtype=torch.float32)
    img_lo = torch.zeros(1280, dtype=torch.float32)
    img_hi[off.reshape(-1)] = hi.reshape(-1)
    img_lo[off.reshape(-1)] = lo.reshape(-1)
    return img_hi.to(device), img_lo.to(device)


def pack_xyz4(xyz_cloud: Cloud, device, B, N):
    """[B,N,4] float32 (x, y, z, 0) copy of a cloud in any layout."""
    out = torch.empty(B, N, 4, dtype=torch.float32, device=device)
    check(lib().dvcp_pack_xyz4(xyz_cloud, B, N, ptr(out), stream_ptr(device)), "dvcp_pack_xyz4")
    _count(1)
    return out


def dfe_tgt_tc(cand, tgt_cloud, tgt_feat, knn_dist, knn_idx32, B, N, b_hi, b_lo, quirks, feature_major_c=0):
    """-> [B,Q,32]; feature_major_c = C > 0: [B,Q/C,32,C] (every key-point's block feature-major)."""
    require_cuda(cand, tgt_feat, knn_dist, knn_idx32, b_hi, b_lo)
    Q = knn_dist.shape[1]
    out = (torch.empty(B, Q, 32, dtype=torch.float32, device=cand.device) if feature_major_c == 0 else
           torch.empty(B, Q // feature_major_c, 32, feature_major_c, dtype=torch.float32, device=cand.device))
    check(lib().dvcp_dfe_tgt_tc(ptr(_f32c(cand)), tgt_cloud, ptr(_f32c(tgt_feat)), ptr(knn_dist), ptr(knn_idx32), B, N,
                                Q, ptr(b_hi), ptr(b_lo), quirks, feature_major_c, ptr(out), stream_ptr(cand.device)),
          "dvcp_dfe_tgt_tc")
    _count(1)
    return out


def dfe_dense(X, dfe):
    """X [..., K, 35] float32/float64 -> [..., 32]."""
    require_cuda(X)
    if X.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("DFE input must be float32 or float64")
    X = X.contiguous()
    K = X.shape[-2]
    rows = X.numel() // (K * 35)
    out = torch.empty(*X.shape[:-2], 32, dtype=torch.float32, device=X.device)
    check(lib().dvcp_dfe_dense(ptr(X), 0 if X.dtype == torch.float32 else 1, rows, K, dfe, ptr(out),
                               stream_ptr(X.device)), "dvcp_dfe_dense")
    _count(1)
    return out


CPG_AUTO, CPG_FUSED, CPG_LAYERED, CPG_TC, CPG_TCZ = 0, 1, 2, 3, 4


def cpg(src_dfe, tgt_dfe, layout, cand, G, params, want_logits=False, path=CPG_AUTO):
    """src_dfe [M,32]; tgt_dfe [M,32*C] flat (layout 0: logical [32,C]; 1: [C,32]); cand [M,C,3].
    path: kernel family (CPG_AUTO = the library's choice; tests force the others)."""
    require_cuda(src_dfe, tgt_dfe, cand)
    M = src_dfe.shape[0]
    C = G * G * G
    dev = src_dfe.device
    nbytes = lib().dvcp_cpg_workspace_bytes(M, G)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    vcp = torch.empty(M, 3, dtype=torch.float32, device=dev)
    logits = torch.empty(M, C, dtype=torch.float32, device=dev) if want_logits else None
    check(lib().dvcp_cpg_path(ptr(_f32c(src_dfe)), ptr(_f32c(tgt_dfe)), layout, ptr(_f32c(cand)), M, G, params,
                              ptr(vcp), ptr(logits), ptr(ws), nbytes, path, stream_ptr(dev)), "dvcp_cpg")
    tc = path in (CPG_TC, CPG_TCZ) or (path == CPG_AUTO and layout == 0 and 2 <= G <= 11)
    _count(2 if tc else (1 if G <= 11 and path != CPG_LAYERED else 5))   # TC: weight image + kernel; fused: 1; layered: 5
    return vcp, logits


def kabsch(x, y, quirks=_lib.QUIRKS_REFERENCE, weights=None):
    """x, y [B,3,n] float32/float64 -> R [B,3,3], t [B,3,1] float64. weights [B,n] (optional):
    weighted solve; quirks without QUIRK_NO_REFLECTION_FIX: det R = +1."""
    require_cuda(x, y, weights)
    if x.dtype != y.dtype or x.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("kabsch: x and y must share a float32/float64 dtype")
    B, _, n = x.shape
    R = torch.empty(B, 3, 3, dtype=torch.float64, device=x.device)
    t = torch.empty(B, 3, 1, dtype=torch.float64, device=x.device)
    if weights is not None:
        weights = weights.to(torch.float64).reshape(B, n).contiguous()
    check(lib().dvcp_kabsch(ptr(x.contiguous()), ptr(y.contiguous()), 0 if x.dtype == torch.float32 else 1,
                            ptr(weights), B, n, quirks, ptr(R), ptr(t), stream_ptr(x.device)), "dvcp_kabsch")
    _count(1)
    return R, t


def kabsch_refine(x, y_pred, R_true, t_true, inlier_ratio=0.8, want_first=False, quirks=_lib.QUIRKS_REFERENCE,
                  want_inliers=False):
    require_cuda(x, y_pred, R_true, t_true)
    B, _, n = x.shape
    dev = x.device
    x = x.double().contiguous()
    y_pred = y_pred.double().contiguous()
    R_true = R_true.double().contiguous()
    t_true = t_true.double().reshape(B, 3, -1)[:, :, 0].contiguous()
    keep = int(n * inlier_ratio)
    R2 = torch.empty(B, 3, 3, dtype=torch.float64, device=dev)
    t2 = torch.empty(B, 3, 1, dtype=torch.float64, device=dev)
    R1 = torch.empty(B, 3, 3, dtype=torch.float64, device=dev) if want_first else None
    t1 = torch.empty(B, 3, 1, dtype=torch.float64, device=dev) if want_first else None
    inl = torch.empty(B, keep, dtype=torch.int64, device=dev) if want_inliers else None
    check(lib().dvcp_kabsch_refine(ptr(x), ptr(y_pred), ptr(R_true), ptr(t_true), B, n, keep, quirks, ptr(R2), ptr(t2),
                                   ptr(R1), ptr(t1), ptr(inl), stream_ptr(dev)), "dvcp_kabsch_refine")
    _count(1)
    if want_inliers:
        return R2, t2, R1, t1, inl
    return R2, t2, R1, t1


def pose_from_forward(src_keypts, tgt_vcp, R_true, t_true, inlier_ratio=0.8, quirks=_lib.QUIRKS_REFERENCE):
    """svd_optimization on the forward's [B,n,3] float32 outputs, read in place (no permute / cast copies)."""
    require_cuda(src_keypts, tgt_vcp, R_true, t_true)
    B, n, _ = src_keypts.shape
    dev = src_keypts.device
    if src_keypts.dtype != torch.float32 or tgt_vcp.dtype != torch.float32:
        raise RuntimeError("pose_from_forward: float32 key-points expected")
    kp, vcp = src_keypts.contiguous(), tgt_vcp.contiguous()     # no-ops for the forward's outputs
    Rt = R_true if R_true.dtype == torch.float64 and R_true.is_contiguous() else R_true.double().contiguous()
    tt = t_true.reshape(B, 3, -1)[:, :, 0] if t_true.dim() == 3 else t_true.reshape(B, 3)
    tt = tt if tt.dtype == torch.float64 and tt.is_contiguous() else tt.double().contiguous()
    R2 = torch.empty(B, 3, 3, dtype=torch.float64, device=dev)
    t2 = torch.empty(B, 3, 1, dtype=torch.float64, device=dev)
    check(lib().dvcp_pose_from_forward(ptr(kp), ptr(vcp), ptr(Rt), ptr(tt), B, n, int(n * inlier_ratio), quirks,
                                       ptr(R2), ptr(t2), stream_ptr(dev)), "dvcp_pose_from_forward")
    _count(1)
    return R2, t2
