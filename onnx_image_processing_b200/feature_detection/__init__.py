from .shi_tomasi_bad import ShiTomasiBADDetector
from .shi_tomasi_bad_sinkhorn import ShiTomasiBADSinkhornMatcher
from .shi_tomasi_sparse_bad_sinkhorn import ShiTomasiSparseBADSinkhornMatcher
from .shi_tomasi_angle import ShiTomasiWithAngle, ShiTomasiAngleSparseBAD, ShiTomasiAngleSparseBADDetector
from .shi_tomasi_angle_sparse_bad_sinkhorn import (ShiTomasiAngleSparseBADSinkhornMatcher,
                                                   ShiTomasiAngleSparseBADSinkhornMatcherWithFilters)
from .shi_tomasi_angle_sparse_bad_sinkhorn_essential_matrix import ShiTomasiAngleSparseBADSinkhornWithEssentialMatrix
from .match_extraction_wrapper import MatchExtractionWrapper
from .akaze_sparse_bad_sinkhorn import AKAZESparseBADSinkhornMatcher

__all__ = [
    "ShiTomasiBADDetector", "ShiTomasiBADSinkhornMatcher", "ShiTomasiSparseBADSinkhornMatcher",
    "ShiTomasiWithAngle", "ShiTomasiAngleSparseBAD", "ShiTomasiAngleSparseBADDetector",
    "ShiTomasiAngleSparseBADSinkhornMatcher", "ShiTomasiAngleSparseBADSinkhornMatcherWithFilters",
    "MatchExtractionWrapper", "ShiTomasiAngleSparseBADSinkhornWithEssentialMatrix", "AKAZESparseBADSinkhornMatcher",
]
