"""MutualNearestNeighborMatcher: drop-in for pytorch_model/matching/match_extraction.py:11-184."""
import torch
from torch import nn

from .. import _ops


class MutualNearestNeighborMatcher(nn.Module):
    """Mutual nearest-neighbour matches from a Sinkhorn matrix: P (B,N+1,M+1), keypoints (B,N,2),(B,M,2) ->
    (matched_kpts1 (B,max_matches,2), matched_kpts2, scores (B,max_matches), valid_mask (B,max_matches) bool)."""

    def __init__(self, max_matches: int = 100, threshold: float = 0.1) -> None:
        super().__init__()
        self.max_matches = max_matches
        self.threshold = threshold

    def forward(self, P: torch.Tensor, keypoints1: torch.Tensor, keypoints2: torch.Tensor):
        return _ops.mutual_matches(P, keypoints1, keypoints2, int(self.max_matches), float(self.threshold))
