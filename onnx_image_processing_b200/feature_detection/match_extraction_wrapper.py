"""MatchExtractionWrapper: drop-in for pytorch_model/feature_detection/match_extraction_wrapper.py:14-113."""
import torch
from torch import nn

from .. import _ops
from ..matching.match_extraction import MutualNearestNeighborMatcher


class MatchExtractionWrapper(nn.Module):
    """Adds mutual nearest-neighbour match extraction behind any matcher that returns
    (keypoints1, keypoints2, matching_probs, ...)."""

    def __init__(self, feature_matcher: nn.Module, max_matches: int = 100, match_threshold: float = 0.1) -> None:
        super().__init__()
        self.feature_matcher = feature_matcher
        self.match_extractor = MutualNearestNeighborMatcher(max_matches=max_matches, threshold=match_threshold)

    def forward(self, image1: torch.Tensor, image2: torch.Tensor):
        fm, mx = self.feature_matcher, self.match_extractor
        if hasattr(fm, "_match_args"):
            # one of this package's unified matchers: ONE C call, the mutual nearest-neighbour extraction (and the outlier
            # filters of a ...WithFilters matcher) runs in the Sinkhorn kernel's epilogue and the (K+1)^2 matrix is never
            # written (match_extraction_wrapper.py:100-113 computes the same from a stored P)
            filt = fm._filter_args() if hasattr(fm, "_filter_args") else (False, -1.0, -1.0)
            out = _ops.match_pairs_ex(image1, image2, *fm._match_args(), False, False, *filt, int(mx.max_matches),
                                      float(mx.threshold))
            return out[6], out[7], out[8], out[9]
        outputs = fm(image1, image2)
        return mx(outputs[2], outputs[0], outputs[1])
