from .sinkhorn import SinkhornMatcher, SinkhornMatcherWithScores, SinkhornMatcherWithFilters
from .match_extraction import MutualNearestNeighborMatcher

__all__ = ["SinkhornMatcher", "SinkhornMatcherWithScores", "SinkhornMatcherWithFilters", "MutualNearestNeighborMatcher"]
