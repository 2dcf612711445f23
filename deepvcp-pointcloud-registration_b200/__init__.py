"""B200-native DeepVCP registration hot path behind the reference's module API.

    from importlib import import_module
    dv = import_module("deepvcp-pointcloud-registration_b200")
    model = dv.DeepVCP(use_normal=False, npoint=16384, r=2.0, s=0.4).cuda().eval()
    src_keypts, tgt_vcp = model(src, tgt, R_init, t_init)
    R, t = dv.pose_from_forward(src_keypts, tgt_vcp, R_gt, t_gt)

Everything numeric runs in libdvcp_b200.so (include/dvcp_b200.h); see DESIGN.md.
"""
from . import KITTIDataset, ModelNet40Dataset, functional, metrics, pipeline, sharding, synthetic, training              # noqa: F401
from ._lib import (QUIRK_COST_VOLUME_RESHAPE, QUIRK_FPS_ORDER_MISMATCH, QUIRK_IGNORE_T_INIT, QUIRK_KEYPOINT_VIEW,  # noqa: F401
                   QUIRK_NO_REFLECTION_FIX, QUIRK_PER_FEATURE_WEIGHT, QUIRKS_INTENDED, QUIRKS_REFERENCE, lib, lib_path)
from .build import build                                               # noqa: F401
from .cpg import cpg                                                   # noqa: F401
from .deep_feat_embedding import feat_embedding_layer                  # noqa: F401
from .deep_feat_extraction import feat_extraction_layer                # noqa: F401
from .deepVCP import DeepVCP                                           # noqa: F401
from .deepVCP_loss import deepVCP_loss, get_rigid_transform, pose_from_forward, svd_optimization  # noqa: F401
from .get_cat_feat_src import Get_Cat_Feat_Src                         # noqa: F401
from .get_cat_feat_tgt import Get_Cat_Feat_Tgt                         # noqa: F401
from .pipeline import GraphedRegistration, StreamedRegistration        # noqa: F401
from .knn_cuda import KNN                                              # noqa: F401
from .pointnet2_utils import (PointNetSetAbstraction, farthest_point_sample, index_points,  # noqa: F401
                              query_ball_point, sample_and_group, square_distance)
from .voxelize import voxelize                                         # noqa: F401
from .weighting_layer import weighting_layer                           # noqa: F401
