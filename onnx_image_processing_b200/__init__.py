"""B200-native Shi-Tomasi + BAD + Sinkhorn feature matching behind the nn.Module API of
fateshelled/onnx_image_processing (pytorch_model/{detector,descriptor,orientation,matching,utils,
feature_detection}).  All arithmetic runs in hand-written sm_100a CUDA kernels in
_lib/libom_b200.so, reached through the C ABI of include/om_b200.h.  No CPU path exists.
"""
from . import _native  # noqa: F401
from .detector import ShiTomasiScore, AKAZE
from .descriptor import BADDescriptor, SparseBAD
from .orientation import AngleEstimator
from .matching import SinkhornMatcher, SinkhornMatcherWithScores, SinkhornMatcherWithFilters, MutualNearestNeighborMatcher
from .utils import apply_nms_maxpool, select_topk_keypoints
from .geometry import EssentialMatrixEstimator
from .feature_detection import (
    ShiTomasiBADDetector,
    ShiTomasiBADSinkhornMatcher,
    ShiTomasiSparseBADSinkhornMatcher,
    ShiTomasiWithAngle,
    ShiTomasiAngleSparseBAD,
    ShiTomasiAngleSparseBADDetector,
    ShiTomasiAngleSparseBADSinkhornMatcher,
    ShiTomasiAngleSparseBADSinkhornMatcherWithFilters,
    MatchExtractionWrapper,
    ShiTomasiAngleSparseBADSinkhornWithEssentialMatrix,
    AKAZESparseBADSinkhornMatcher,
)
from .ingest import FrameIngest, load_image_from_array

__all__ = [
    "ShiTomasiScore", "BADDescriptor", "SparseBAD", "AngleEstimator", "SinkhornMatcher",
    "SinkhornMatcherWithScores", "apply_nms_maxpool", "select_topk_keypoints", "ShiTomasiBADDetector",
    "ShiTomasiBADSinkhornMatcher", "ShiTomasiSparseBADSinkhornMatcher", "ShiTomasiWithAngle",
    "ShiTomasiAngleSparseBAD", "ShiTomasiAngleSparseBADDetector", "ShiTomasiAngleSparseBADSinkhornMatcher",
    "MutualNearestNeighborMatcher", "MatchExtractionWrapper", "SinkhornMatcherWithFilters",
    "ShiTomasiAngleSparseBADSinkhornMatcherWithFilters", "EssentialMatrixEstimator",
    "ShiTomasiAngleSparseBADSinkhornWithEssentialMatrix", "AKAZE", "AKAZESparseBADSinkhornMatcher", "FrameIngest",
    "load_image_from_array",
]
