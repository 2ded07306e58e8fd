"""SinkhornMatcher: drop-in for pytorch_model/matching/sinkhorn.py:28-259."""
import torch
from torch import nn

from .. import _ops


class SinkhornMatcher(nn.Module):
    """Log-domain Sinkhorn with dustbins: (B,N,D),(B,M,D) -> (B,N+1,M+1) match probabilities."""

    def __init__(self, iterations: int = 20, epsilon: float = 1.0, unused_score: float = 1.0,
                 distance_type: str = "l2") -> None:
        super().__init__()
        if iterations <= 0:
            raise ValueError(f"iterations must be positive, got {iterations}")
        if epsilon <= 0:
            raise ValueError(f"epsilon must be positive, got {epsilon}")
        self.iterations = iterations
        self.epsilon = epsilon
        self.unused_score = unused_score
        self.distance_type = distance_type.lower()
        if self.distance_type not in ("l1", "l2"):
            raise ValueError(f"distance_type must be 'l1' or 'l2', got {distance_type}")

    def forward(self, desc1: torch.Tensor, desc2: torch.Tensor) -> torch.Tensor:
        return _ops.sinkhorn(desc1, desc2, self.iterations, float(self.epsilon), float(self.unused_score),
                             self.distance_type == "l1")


class SinkhornMatcherWithScores(SinkhornMatcher):
    """Also returns the best non-dustbin probability per row and per column (sinkhorn.py:211-259)."""

    def forward(self, desc1: torch.Tensor, desc2: torch.Tensor):
        out = _ops.sinkhorn_ex(desc1, desc2, self.iterations, float(self.epsilon), float(self.unused_score),
                               self.distance_type == "l1", True, True, False, -1.0, -1.0, None, None, 0, 0.0)
        return out[0], out[1], out[2]


class SinkhornMatcherWithFilters(SinkhornMatcher):
    """Sinkhorn + probability-ratio and dustbin-margin outlier filters (sinkhorn.py:262-465): returns
    (P with rejected rows sent to the dustbin, valid_mask (B,N) bool)."""

    def __init__(self, iterations: int = 20, epsilon: float = 1.0, unused_score: float = 1.0, distance_type: str = "l2",
                 ratio_threshold: float = None, dustbin_margin: float = None) -> None:
        super().__init__(iterations, epsilon, unused_score, distance_type)
        self.ratio_threshold = ratio_threshold if ratio_threshold is not None else -1.0
        self.dustbin_margin = dustbin_margin if dustbin_margin is not None else -1.0

    def forward(self, desc1: torch.Tensor, desc2: torch.Tensor):
        out = _ops.sinkhorn_ex(desc1, desc2, self.iterations, float(self.epsilon), float(self.unused_score),
                               self.distance_type == "l1", True, False, True, float(self.ratio_threshold),
                               float(self.dustbin_margin), None, None, 0, 0.0)
        return out[0], out[3]
