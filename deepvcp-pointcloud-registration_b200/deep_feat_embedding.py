"""feat_embedding_layer -- deep_feat_embedding.py:13-61 of the reference: Linear
35->32->32->32 (no activations) and a max over the neighbour axis."""
import torch.nn as nn

from . import functional as F_
from ._lib import dfe_params


class feat_embedding_layer(nn.Module):
    def __init__(self, K_nsample=32):
        super().__init__()
        self.K_nsample = 32
        self.fc1 = nn.Linear(35, 32, True)
        self.fc2 = nn.Linear(32, 32, True)
        self.fc3 = nn.Linear(32, 32, True)
        self._cache = None

    def params(self):
        ps = [self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, self.fc3.weight, self.fc3.bias]
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if self._cache is None or self._cache[0] != key:
            ts = [p.detach().float().contiguous() for p in ps]
            self._cache = (key, ts, dfe_params(*ts))
        return self._cache[2]

    def tc_operand(self):
        """(b_hi, b_lo) operand images of the collapsed affine map for the tcgen05 kernel."""
        ps = [self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, self.fc3.weight, self.fc3.bias]
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if getattr(self, "_tc_cache", None) is None or self._tc_cache[0] != key:
            self._tc_cache = (key, F_.dfe_tc_operand(*ps, device=self.fc1.weight.device))
        return self._tc_cache[1]

    def forward(self, X, src=True):
        """src: [B,N,K,35] -> [B,N,32]; tgt: [B,N,C,K,35] -> [B,N,C,32]."""
        return F_.dfe_dense(X, self.params())
