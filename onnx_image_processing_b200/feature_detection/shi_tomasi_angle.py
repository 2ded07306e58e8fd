"""Drop-ins for pytorch_model/feature_detection/shi_tomasi_angle.py: ShiTomasiWithAngle :23-98,
ShiTomasiAngleSparseBAD :101-243, ShiTomasiAngleSparseBADDetector :246-356."""
import torch
from torch import nn

from .. import _ops
from ..descriptor.bad import SparseBAD
from ..detector.shi_tomasi import ShiTomasiScore
from ..orientation.angle_estimation import AngleEstimator


class ShiTomasiWithAngle(nn.Module):
    """(N,1,H,W) -> (score map, orientation map), both (N,1,H,W)."""

    def __init__(self, block_size: int = 5, sobel_size: int = 3, patch_size: int = 15, sigma: float = 2.5):
        super().__init__()
        self.shi_tomasi = ShiTomasiScore(block_size=block_size, sobel_size=sobel_size)
        self.angle_estimator = AngleEstimator(patch_size=patch_size, sigma=sigma)

    def forward(self, image: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        return self.shi_tomasi(image), self.angle_estimator(image)


class ShiTomasiAngleSparseBAD(nn.Module):
    """Score + orientation maps and rotation-aware sparse BAD at caller-provided keypoints."""

    def __init__(self, block_size: int = 5, patch_size: int = 15, sigma: float = 2.5, num_pairs: int = 256,
                 binarize: bool = False, soft_binarize: bool = True, temperature: float = 10.0,
                 normalize_descriptors: bool = True, sampling_mode: str = "nearest"):
        super().__init__()
        self.detector = ShiTomasiWithAngle(block_size=block_size, patch_size=patch_size, sigma=sigma)
        self.descriptor = SparseBAD(num_pairs=num_pairs, binarize=binarize, soft_binarize=soft_binarize,
                                    temperature=temperature, normalize_descriptors=normalize_descriptors,
                                    sampling_mode=sampling_mode)

    def detect_and_orient(self, image: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        return self.detector(image)

    def describe(self, image: torch.Tensor, keypoints: torch.Tensor, orientation: torch.Tensor) -> torch.Tensor:
        return self.descriptor(image, keypoints, orientation)

    def forward(self, image: torch.Tensor, keypoints: torch.Tensor):
        scores, angles = self.detect_and_orient(image)
        return scores, angles, self.describe(image, keypoints, angles)


class ShiTomasiAngleSparseBADDetector(nn.Module):
    """Single image -> (keypoints (B,K,2), scores (B,K), descriptors (B,K,num_pairs)); no border
    margin (reference :349-351).  The orientation is evaluated at the K keypoints only."""

    def __init__(self, max_keypoints: int, block_size: int = 5, patch_size: int = 15, sigma: float = 2.5,
                 num_pairs: int = 256, binarize: bool = False, soft_binarize: bool = True, temperature: float = 10.0,
                 normalize_descriptors: bool = True, sampling_mode: str = "nearest", nms_radius: int = 3,
                 score_threshold: float = 0.0) -> None:
        super().__init__()
        self.max_keypoints = max_keypoints
        self.nms_radius = nms_radius
        self.score_threshold = score_threshold
        self.model = ShiTomasiAngleSparseBAD(block_size=block_size, patch_size=patch_size, sigma=sigma,
                                             num_pairs=num_pairs, binarize=binarize, soft_binarize=soft_binarize,
                                             temperature=temperature, normalize_descriptors=normalize_descriptors,
                                             sampling_mode=sampling_mode)

    def forward(self, image: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        det, d = self.model.detector, self.model.descriptor
        kpts, scores = _ops.detect(image, int(self.max_keypoints), det.shi_tomasi.block_size, int(self.nms_radius),
                                   float(self.score_threshold), 0)
        desc = _ops.sparse_bad(image, kpts, d._pair_table, d._mode(), float(d.temperature),
                               bool(d.normalize_descriptors), _ops.sampling_code(d.sampling_mode), _ops.THETA_MOMENTS,
                               None, det.angle_estimator.moment_kernels)
        return kpts, scores, desc
