"""ShiTomasiAngleSparseBADSinkhornMatcher: drop-in for
pytorch_model/feature_detection/shi_tomasi_angle_sparse_bad_sinkhorn.py:26-180."""
import torch
from torch import nn

from .. import _ops
from ..descriptor.bad import SparseBAD
from ..matching.sinkhorn import SinkhornMatcher, SinkhornMatcherWithFilters
from .shi_tomasi_angle import ShiTomasiWithAngle


class ShiTomasiAngleSparseBADSinkhornMatcher(nn.Module):
    """Rotation-invariant matcher: Shi-Tomasi(block 5) + moment orientation + rotated sparse BAD +
    Sinkhorn.  The 15x15 moment convolution is evaluated at the K keypoints only, inside the
    descriptor kernel, instead of over the whole image."""

    def __init__(self, max_keypoints: int, block_size: int = 5, patch_size: int = 15, sigma: float = 2.5,
                 num_pairs: int = 256, binarize: bool = False, soft_binarize: bool = True, temperature: float = 10.0,
                 sinkhorn_iterations: int = 20, epsilon: float = 1.0, unused_score: float = 1.0,
                 distance_type: str = "l2", nms_radius: int = 3, score_threshold: float = 0.0,
                 normalize_descriptors: bool = True, sampling_mode: str = "nearest",
                 border_margin: int | None = None) -> None:
        super().__init__()
        self.max_keypoints = max_keypoints
        self.nms_radius = nms_radius
        self.score_threshold = score_threshold
        self.detector = ShiTomasiWithAngle(block_size=block_size, patch_size=patch_size, sigma=sigma)
        self.descriptor = SparseBAD(num_pairs=num_pairs, binarize=binarize, soft_binarize=soft_binarize,
                                    temperature=temperature, normalize_descriptors=normalize_descriptors,
                                    sampling_mode=sampling_mode)
        self.border_margin = self.descriptor.max_radius if border_margin is None else border_margin
        self.matcher = SinkhornMatcher(iterations=sinkhorn_iterations, epsilon=epsilon, unused_score=unused_score,
                                       distance_type=distance_type)

    def _match_args(self):
        d, m = self.descriptor, self.matcher
        return (d._pair_table, self.detector.angle_estimator.moment_kernels,
                _ops.MATCH_ANGLE, int(self.max_keypoints), self.detector.shi_tomasi.block_size,
                int(self.nms_radius), int(self.border_margin), float(self.score_threshold), d._mode(),
                float(d.temperature), bool(d.normalize_descriptors),
                _ops.sampling_code(d.sampling_mode), m.iterations, float(m.epsilon),
                float(m.unused_score), m.distance_type == "l1")

    def match(self, image1: torch.Tensor, image2: torch.Tensor):
        return _ops.match_pairs(image1, image2, *self._match_args())

    def forward(self, image1: torch.Tensor, image2: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        k1, k2, probs, _, _ = self.match(image1, image2)
        return k1, k2, probs


class ShiTomasiAngleSparseBADSinkhornMatcherWithFilters(ShiTomasiAngleSparseBADSinkhornMatcher):
    """The rotation-invariant matcher with the outlier filters built into the matching step
    (shi_tomasi_angle_sparse_bad_sinkhorn.py:183-340): returns (kpts1, kpts2, filtered P, valid_mask)."""

    def __init__(self, max_keypoints: int, block_size: int = 5, patch_size: int = 15, sigma: float = 2.5,
                 num_pairs: int = 256, binarize: bool = False, soft_binarize: bool = True, temperature: float = 10.0,
                 sinkhorn_iterations: int = 20, epsilon: float = 1.0, unused_score: float = 1.0,
                 distance_type: str = "l2", ratio_threshold: float = None, dustbin_margin: float = None,
                 nms_radius: int = 3, score_threshold: float = 0.0, normalize_descriptors: bool = True,
                 sampling_mode: str = "nearest", border_margin: int | None = None) -> None:
        super().__init__(max_keypoints, block_size, patch_size, sigma, num_pairs, binarize, soft_binarize, temperature,
                         sinkhorn_iterations, epsilon, unused_score, distance_type, nms_radius, score_threshold,
                         normalize_descriptors, sampling_mode, border_margin)
        self.matcher = SinkhornMatcherWithFilters(iterations=sinkhorn_iterations, epsilon=epsilon, unused_score=unused_score,
                                                  distance_type=distance_type, ratio_threshold=ratio_threshold,
                                                  dustbin_margin=dustbin_margin)

    def _filter_args(self):
        return True, float(self.matcher.ratio_threshold), float(self.matcher.dustbin_margin)

    def forward(self, image1: torch.Tensor, image2: torch.Tensor):
        # one C call: the filters run in the Sinkhorn kernel's epilogue on the probabilities it holds in registers
        out = _ops.match_pairs_ex(image1, image2, *self._match_args(), True, False, *self._filter_args(), 0, 0.0)
        return out[0], out[1], out[2], out[5]
