"""ctypes binding of libdvcp_b200.so (include/dvcp_b200.h) for torch tensors.

PyTorch is plumbing here: it owns device memory and the stream; every compute
call goes through the C ABI with raw device pointers. There is no CPU or
PyTorch fallback: if the library cannot be loaded, or a tensor is not on a CUDA
device, the call raises.
"""
import ctypes
import os

import torch

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_i64, c_i32, c_f32, c_f64, c_vp = (ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_double,
                                    ctypes.c_void_p)


class Cloud(ctypes.Structure):
    _fields_ = [("base", c_vp), ("bstride", c_i64), ("pstride", c_i64), ("cstride", c_i64)]


class CloudIndex(ctypes.Structure):
    _fields_ = [("sorted_pt", c_vp), ("bucket_box", c_vp), ("cap", c_i32)]


class MlpLayer(ctypes.Structure):
    _fields_ = [("W", c_vp), ("b", c_vp), ("alpha", c_vp), ("beta", c_vp), ("in_ch", c_i32), ("out_ch", c_i32)]


class DfeParams(ctypes.Structure):
    _fields_ = [(n, c_vp) for n in ("W1", "b1", "W2", "b2", "W3", "b3")]


class CpgParams(ctypes.Structure):
    _fields_ = [(n, c_vp) for n in ("w1", "b1", "w2", "b2", "w3", "b3")]


# name -> (restype, argtypes); mirrors include/dvcp_b200.h one to one
SIGNATURES = {
    "dvcp_abi_version": (c_i32, []),
    "dvcp_error_string": (ctypes.c_char_p, [c_i32]),
    "dvcp_fps": (c_i32, [Cloud, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, CloudIndex, c_vp]),
    "dvcp_fps_indexed": (c_i32, [Cloud, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, CloudIndex, c_i32, c_vp]),
    "dvcp_index_capacity": (c_i32, [c_i32]),
    "dvcp_build_index": (c_i32, [Cloud, c_i32, c_i32, CloudIndex, c_vp]),
    "dvcp_index_capacity_any": (c_i32, [c_i32]),
    "dvcp_build_index_workspace_bytes": (c_i64, [c_i32, c_i32]),
    "dvcp_build_index_ws": (c_i32, [Cloud, c_i32, c_i32, CloudIndex, c_vp, c_i64, c_vp]),
    "dvcp_fps_plain": (c_i32, [Cloud, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "dvcp_square_distance": (c_i32, [Cloud, Cloud, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "dvcp_ball_query": (c_i32, [Cloud, Cloud, c_i32, c_i32, c_i32, c_f32, c_i32, c_vp, c_vp]),
    "dvcp_index_points": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i64, c_vp, c_vp]),
    "dvcp_index_points_i32": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i64, c_vp, c_vp]),
    "dvcp_sa_layer": (c_i32, [Cloud, Cloud, c_i32, c_vp, c_i32, c_i32, c_i32, c_f32, c_i32,
                              ctypes.POINTER(MlpLayer), c_i32, CloudIndex, c_vp, c_vp, c_vp, c_vp]),
    "dvcp_sa_layer_all": (c_i32, [Cloud, Cloud, c_i32, c_vp, c_i32, c_i32, c_f32, c_i32, ctypes.POINTER(MlpLayer), c_i32,
                                  CloudIndex, c_vp, c_vp, c_vp]),
    "dvcp_linear_rows": (c_i32, [c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "dvcp_weighting_scores": (c_i32, [c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "dvcp_topk": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "dvcp_keypoint_stage": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_vp, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp, c_i64,
                                    c_f32, c_i32, DfeParams, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "dvcp_square_distance_f64": (c_i32, [Cloud, Cloud, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "dvcp_ball_query_f64": (c_i32, [Cloud, Cloud, c_i32, c_i32, c_i32, c_f64, c_i32, c_vp, c_vp]),
    "dvcp_sa_layer_f64": (c_i32, [Cloud, Cloud, c_i32, c_vp, c_i32, c_i32, c_i32, c_f64, c_i32,
                                  ctypes.POINTER(MlpLayer), c_i32, c_vp, c_vp, c_vp]),
    "dvcp_keypoint_stage_f64": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_vp, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp, c_i64,
                                        c_f64, c_i32, DfeParams, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "dvcp_grid_size": (c_i32, [c_f64, c_f64]),
    "dvcp_candidates": (c_i32, [c_vp, c_i64, c_f64, c_f64, c_i32, c_vp, c_vp]),
    "dvcp_knn": (c_i32, [Cloud, c_vp, c_i32, c_i32, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "dvcp_knn_indexed": (c_i32, [CloudIndex, c_vp, c_i32, c_i32, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp,
                                 c_vp]),
    "dvcp_knn_groups": (c_i32, [CloudIndex, c_vp, c_i32, c_i32, c_i64, c_i32, c_i32, c_i32, c_f32, c_i32, c_vp, c_vp,
                                c_vp, c_vp, c_vp, c_i64, c_vp]),
    "dvcp_knn_groups_workspace_bytes": (c_i64, [c_i32, c_i64, c_i32]),
    "dvcp_dfe_tgt_fused": (c_i32, [c_vp, Cloud, c_vp, c_vp, c_vp, c_i32, c_i32, c_i64, DfeParams, c_i32, c_vp,
                                   c_vp]),
    "dvcp_dfe_tgt_backward": (c_i32, [c_vp, Cloud, c_vp, c_vp, c_vp, c_i32, c_i32, c_i64, DfeParams, c_vp, c_i32, c_vp, c_vp,
                                      c_vp, c_vp, c_vp]),
    "dvcp_dfe_tc_b_floats": (c_i32, []),
    "dvcp_dfe_tc_b_offset": (c_i32, [c_i32, c_i32]),
    "dvcp_dfe_tgt_tc": (c_i32, [c_vp, Cloud, c_vp, c_vp, c_vp, c_i32, c_i32, c_i64, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp]),
    "dvcp_cpg_tc_image_bytes": (c_i64, []),
    "dvcp_dfe_dense": (c_i32, [c_vp, c_i32, c_i64, c_i32, DfeParams, c_vp, c_vp]),
    "dvcp_cpg_workspace_bytes": (c_i64, [c_i64, c_i32]),
    "dvcp_cpg": (c_i32, [c_vp, c_vp, c_i32, c_vp, c_i64, c_i32, CpgParams, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "dvcp_cpg_path": (c_i32, [c_vp, c_vp, c_i32, c_vp, c_i64, c_i32, CpgParams, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp]),
    "dvcp_pose_from_forward": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "dvcp_pack_xyz4": (c_i32, [Cloud, c_i32, c_i32, c_vp, c_vp]),
    "dvcp_ingest_kitti": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "dvcp_ingest_modelnet": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "dvcp_voxel_filter_workspace_bytes": (c_i64, [c_i64]),
    "dvcp_voxel_grid_filter": (c_i32, [c_vp, c_i32, c_i32, c_i64, c_f32, c_f32, c_f32, c_f32, c_i32, c_vp, c_i64, c_vp, c_vp,
                                       c_vp, c_vp]),
    "dvcp_kabsch": (c_i32, [c_vp, c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "dvcp_kabsch_refine": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp,
                                   c_vp, c_vp]),
}

# A set bit replicates the reference; a clear bit selects the semantics its code intends
QUIRK_KEYPOINT_VIEW = 1         # Q3: gather result re-read row-major (deepVCP.py:46)
QUIRK_PER_FEATURE_WEIGHT = 2    # Q7: distance weight indexed by feature (get_cat_feat_tgt.py:65,92)
QUIRK_COST_VOLUME_RESHAPE = 4   # Q4: (feature, candidate) axes scrambled by the reshape (deepVCP.py:106, cpg.py:34)
QUIRK_IGNORE_T_INIT = 8         # Q6: t_init never added (deepVCP.py:86-91)
QUIRK_NO_REFLECTION_FIX = 16    # Q10: R = V U^T may have det = -1 (deepVCP_loss.py:36-40)
QUIRK_FPS_ORDER_MISMATCH = 32   # Q5: rows of the FPS-ordered feature tables addressed with original-order / key-point-local indices
                                #     (deepVCP.py:35,46,61; get_cat_feat_tgt.py:85) -- host-side wiring, no kernel reads this bit
QUIRKS_REFERENCE = 63
QUIRKS_INTENDED = 0


def lib_path() -> str:
    return _build.LIB


def lib():
    """Loads (building first if the sources are newer) the C-ABI library."""
    global _LIB
    if _LIB is None:
        path = _build.LIB
        if _build.needs_build():
            path = _build.build()
        L = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.dvcp_abi_version() != 2:
            raise RuntimeError("libdvcp_b200.so ABI version mismatch")
        _LIB = L
    return _LIB


def check(code: int, what: str):
    if code != 0:
        msg = lib().dvcp_error_string(code).decode()
        raise RuntimeError("%s failed (%d): %s" % (what, code, msg))


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("deepvcp_b200 runs on CUDA (sm_100a) only; got a %s tensor. "
                               "There is no CPU fallback." % t.device)


def ptr(t):
    return None if t is None else t.data_ptr()


def cloud_pm(t: torch.Tensor) -> Cloud:
    """Point-major view [B, N, C>=3] (any strides) -> Cloud."""
    return Cloud(t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))


def cloud_cm(t: torch.Tensor) -> Cloud:
    """Channel-major view [B, C>=3, N] (any strides) -> Cloud."""
    return Cloud(t.data_ptr(), t.stride(0), t.stride(2), t.stride(1))


NULL_CLOUD = Cloud(None, 0, 0, 0)
NULL_INDEX = CloudIndex(None, None, 0)


def dfe_params(w1, b1, w2, b2, w3, b3) -> DfeParams:
    return DfeParams(*(t.data_ptr() for t in (w1, b1, w2, b2, w3, b3)))


def cpg_params(w1, b1, w2, b2, w3, b3) -> CpgParams:
    return CpgParams(*(t.data_ptr() for t in (w1, b1, w2, b2, w3, b3)))
