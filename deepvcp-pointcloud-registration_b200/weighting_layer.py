"""weighting_layer -- weighting_layer.py:8-33 of the reference: per-point MLP
32->16->8->1 (ReLU, ReLU, Softplus) and the indices of the K highest scores,
flattened. Ties are ordered (score desc, index asc); torch.topk leaves them
unspecified."""
import torch.nn as nn

from . import functional as F_


class weighting_layer(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1 = nn.Sequential(nn.Linear(32, 16, True), nn.ReLU())
        self.fc2 = nn.Sequential(nn.Linear(16, 8, True), nn.ReLU())
        self.fc3 = nn.Sequential(nn.Linear(8, 1, True), nn.Softplus())

    def scores(self, X):
        p = [self.fc1[0].weight, self.fc1[0].bias, self.fc2[0].weight, self.fc2[0].bias, self.fc3[0].weight,
             self.fc3[0].bias]
        return F_.weighting_scores(X, *[t.detach().float().contiguous() for t in p])

    def forward(self, X, K=64):
        """X [B,S,32] -> [B*K] int64."""
        return F_.topk(self.scores(X), K).flatten()
