"""MatchExtractionWrapper: drop-in for pytorch_model/feature_detection/match_extraction_wrapper.py:14-113."""
import torch
from torch import nn

from ..matching.match_extraction import MutualNearestNeighborMatcher


class MatchExtractionWrapper(nn.Module):
    """Adds mutual nearest-neighbour match extraction behind any matcher that returns
    (keypoints1, keypoints2, matching_probs, ...)."""

    def __init__(self, feature_matcher: nn.Module, max_matches: int = 100, match_threshold: float = 0.1) -> None:
        super().__init__()
        self.feature_matcher = feature_matcher
        self.match_extractor = MutualNearestNeighborMatcher(max_matches=max_matches, threshold=match_threshold)

    def forward(self, image1: torch.Tensor, image2: torch.Tensor):
        outputs = self.feature_matcher(image1, image2)
        return self.match_extractor(outputs[2], outputs[0], outputs[1])
