"""KNN -- stand-in for the third-party `knn_cuda.KNN` the reference imports
(get_cat_feat_tgt.py:4,45,52; deepVCP_loss.py:3,70,72): same constructor and call
convention, exact search on sm_100a: through the spatial index for clouds of
64..131072 points and at least 256 queries, brute force otherwise; identical results."""
import torch

from . import functional as F_
from ._lib import cloud_cm, cloud_pm, require_cuda


class KNN:
    def __init__(self, k, transpose_mode=False):
        self.k = k
        self.transpose_mode = transpose_mode

    def __call__(self, ref, query):
        """transpose_mode=True: ref [B,N,3], query [B,Q,3] -> dist, idx [B,Q,k];
        False: ref [B,3,N], query [B,3,Q] -> [B,k,Q]."""
        require_cuda(ref, query)
        with torch.no_grad():
            ref = ref.float()
            query = query.float()
            if self.transpose_mode:
                B, N, _ = ref.shape
                rc, q = cloud_pm(ref), query.contiguous()
            else:
                B, _, N = ref.shape
                rc, q = cloud_cm(ref), query.transpose(1, 2).contiguous()
            if F_.SpatialIndex.knn_indexable(N) and q.shape[1] >= 256:
                index = F_.build_index(rc, ref.device, B, N, big=True)
                dist, idx, _ = F_.knn_indexed(index, 0, ref.device, B, N, q, self.k)
            else:
                dist, idx, _ = F_.knn(rc, ref.device, B, N, q, self.k)
            if not self.transpose_mode:
                dist, idx = dist.transpose(1, 2).contiguous(), idx.transpose(1, 2).contiguous()
        return dist, idx
