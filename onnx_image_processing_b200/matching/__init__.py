from .sinkhorn import SinkhornMatcher, SinkhornMatcherWithScores
from .match_extraction import MutualNearestNeighborMatcher

__all__ = ["SinkhornMatcher", "SinkhornMatcherWithScores", "MutualNearestNeighborMatcher"]
