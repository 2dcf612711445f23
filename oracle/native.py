"""ctypes binding of oracle/dvcp_oracle.c (test infrastructure only)."""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdvcp_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "dvcp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        i64, f32, f64, vp = ctypes.c_int64, ctypes.c_float, ctypes.c_double, ctypes.c_void_p
        L.orc_fps_f32.argtypes = [vp, i64, i64, i64, vp]
        L.orc_fps_f64.argtypes = [vp, i64, i64, i64, vp]
        L.orc_square_distance_f32.argtypes = [vp, i64, vp, i64, vp]
        L.orc_ball_query_f32.argtypes = [vp, i64, vp, i64, f32, i64, vp]
        L.orc_ball_query_f64.argtypes = [vp, i64, vp, i64, f64, i64, vp]
        L.orc_square_distance_f64.argtypes = [vp, i64, vp, i64, vp]
        L.orc_knn_f32.argtypes = [vp, i64, vp, i64, i64, vp, vp]
        L.orc_candidates.argtypes = [vp, i64, f64, f64, i64, vp]
        L.orc_grid_size.argtypes = [f64, f64, f64]
        L.orc_grid_size.restype = i64
        L.orc_voxel_filter.argtypes = [vp, i64, i64, i64, f32, f32, f32, f32, i64, vp, vp]
        L.orc_voxel_filter.restype = i64
        for fn in (L.orc_fps_f32, L.orc_fps_f64, L.orc_square_distance_f32,
                   L.orc_ball_query_f32, L.orc_knn_f32, L.orc_candidates, L.orc_ball_query_f64,
                   L.orc_square_distance_f64):
            fn.restype = None
        _lib = L
    return _lib


def _c(t: torch.Tensor, dtype) -> torch.Tensor:
    return t.detach().to("cpu", dtype).contiguous()


def fps(xyz: torch.Tensor, npoint: int, start) -> torch.Tensor:
    """xyz [B,N,3] (f32 or f64), start [B] -> centroids [B,npoint] int64."""
    B, N, _ = xyz.shape
    out = torch.empty(B, npoint, dtype=torch.int64)
    start = [int(s) for s in torch.as_tensor(start).reshape(-1)]
    if xyz.dtype == torch.float64:
        x = _c(xyz, torch.float64)
        fn = lib().orc_fps_f64
    else:
        x = _c(xyz, torch.float32)
        fn = lib().orc_fps_f32
    for b in range(B):
        fn(x[b].data_ptr(), N, npoint, start[b], out[b].data_ptr())
    return out


def square_distance(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    B, S, _ = src.shape
    N = dst.shape[1]
    if src.dtype == torch.float64 or dst.dtype == torch.float64:   # torch promotes: the matmul runs in double
        a, d = _c(src, torch.float64), _c(dst, torch.float64)
        out = torch.empty(B, S, N, dtype=torch.float64)
        for b in range(B):
            lib().orc_square_distance_f64(a[b].data_ptr(), S, d[b].data_ptr(), N, out[b].data_ptr())
        return out
    a, d = _c(src, torch.float32), _c(dst, torch.float32)
    out = torch.empty(B, S, N, dtype=torch.float32)
    for b in range(B):
        lib().orc_square_distance_f32(a[b].data_ptr(), S, d[b].data_ptr(), N, out[b].data_ptr())
    return out


def ball_query(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    out = torch.empty(B, S, nsample, dtype=torch.int64)
    if xyz.dtype == torch.float64 or new_xyz.dtype == torch.float64:
        x, q = _c(xyz, torch.float64), _c(new_xyz, torch.float64)
        for b in range(B):
            lib().orc_ball_query_f64(x[b].data_ptr(), N, q[b].data_ptr(), S, float(radius ** 2), nsample,
                                     out[b].data_ptr())
        return out
    x, q = _c(xyz, torch.float32), _c(new_xyz, torch.float32)
    r2 = float(np.float32(radius ** 2))
    for b in range(B):
        lib().orc_ball_query_f32(x[b].data_ptr(), N, q[b].data_ptr(), S, r2, nsample, out[b].data_ptr())
    return out


def knn(ref: torch.Tensor, query: torch.Tensor, k: int):
    """ref [B,N,3], query [B,Q,3] -> dist [B,Q,k] f32 (sqrt), idx [B,Q,k] int64."""
    B, N, _ = ref.shape
    Q = query.shape[1]
    r, q = _c(ref, torch.float32), _c(query, torch.float32)
    dist = torch.empty(B, Q, k, dtype=torch.float32)
    idx = torch.empty(B, Q, k, dtype=torch.int64)
    for b in range(B):
        lib().orc_knn_f32(r[b].data_ptr(), N, q[b].data_ptr(), Q, k, dist[b].data_ptr(), idx[b].data_ptr())
    return dist, idx


def grid_size(r: float, s: float, c: float = 0.0) -> int:
    return int(lib().orc_grid_size(float(c), float(r), float(s)))


def candidates(centres: torch.Tensor, r: float, s: float, G: int = None) -> torch.Tensor:
    """centres [B,M,3] float64 -> [B,M,G^3,3] float32."""
    B, M, _ = centres.shape
    c = _c(centres, torch.float64)
    if G is None:
        G = grid_size(r, s, float(c.reshape(-1)[0]))
    out = torch.empty(B, M, G * G * G, 3, dtype=torch.float32)
    lib().orc_candidates(c.data_ptr(), B * M, float(r), float(s), G, out.data_ptr())
    return out


def voxel_filter(pts: torch.Tensor, cell: float, origin=(0.0, 0.0, 0.0), mode: str = "centroid"):
    """pts [M, 3 or 4] float32 -> (out [n, C] float32, counts [n] int32): see orc_voxel_filter."""
    x = _c(pts, torch.float32)
    M, C = x.shape
    out = torch.empty(M, C, dtype=torch.float32)
    cnt = torch.empty(M, dtype=torch.int32)
    n = lib().orc_voxel_filter(x.data_ptr(), M, C, C, float(origin[0]), float(origin[1]), float(origin[2]), float(cell),
                               {"centroid": 0, "first": 1}[mode], out.data_ptr(), cnt.data_ptr())
    return out[:n].clone(), cnt[:n].clone()
