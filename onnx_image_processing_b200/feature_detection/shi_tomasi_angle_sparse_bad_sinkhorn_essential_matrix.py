"""ShiTomasiAngleSparseBADSinkhornWithEssentialMatrix: drop-in for
pytorch_model/feature_detection/shi_tomasi_angle_sparse_bad_sinkhorn_essential_matrix.py:34-361."""
import torch

from ..geometry.essential_matrix_estimator import EssentialMatrixEstimator
from .shi_tomasi_angle_sparse_bad_sinkhorn import ShiTomasiAngleSparseBADSinkhornMatcher


class ShiTomasiAngleSparseBADSinkhornWithEssentialMatrix(ShiTomasiAngleSparseBADSinkhornMatcher):
    """Rotation-invariant matcher + essential matrix from the detected keypoints: returns (keypoints1, keypoints2,
    matching_probs, E).  E is (3, 3) for a batch of one pair, as in the reference (which requires batch 1), and
    (B, 3, 3) for B > 1 pairs."""

    def __init__(self, K: torch.Tensor, max_keypoints: int, block_size: int = 5, patch_size: int = 15, sigma: float = 2.5,
                 num_pairs: int = 256, binarize: bool = False, soft_binarize: bool = True, temperature: float = 10.0,
                 sinkhorn_iterations: int = 20, epsilon: float = 1.0, unused_score: float = 1.0,
                 distance_type: str = "l2", nms_radius: int = 3, score_threshold: float = 0.0,
                 normalize_descriptors: bool = True, sampling_mode: str = "nearest", border_margin: int | None = None,
                 top_k: int = 3, n_iter: int = 30, n_iter_manifold: int = 10) -> None:
        super().__init__(max_keypoints, block_size, patch_size, sigma, num_pairs, binarize, soft_binarize, temperature,
                         sinkhorn_iterations, epsilon, unused_score, distance_type, nms_radius, score_threshold,
                         normalize_descriptors, sampling_mode, border_margin)
        self.top_k = top_k
        self.estimator = EssentialMatrixEstimator(K=K, image_shape=(1, 1), top_k=top_k, n_iter=n_iter,
                                                  n_iter_manifold=n_iter_manifold)          # :160-166
        self.register_buffer("K_inv", torch.linalg.inv(K.float()))                          # :170-172

    def _normalised(self, keypoints_yx: torch.Tensor) -> torch.Tensor:
        """(B, K, 2) (y, x) pixels -> (x, y) normalised image coordinates, :341-352"""
        xy = torch.stack([keypoints_yx[..., 1], keypoints_yx[..., 0]], dim=-1)
        ones = xy.new_ones(*xy.shape[:-1], 1)
        return (torch.cat([xy, ones], dim=-1) @ self.K_inv.T.to(xy))[..., :2].contiguous()

    def forward(self, image1: torch.Tensor, image2: torch.Tensor):
        k1, k2, probs, _, _ = self.match(image1, image2)
        # valid <=> keypoint score > 0 <=> the keypoint is not the (-1, -1) padding (keypoint_utils.py:108-114), :337-338
        E = self.estimator.estimate(probs, self._normalised(k1), self._normalised(k2), k1[..., 0] >= 0, k2[..., 0] >= 0)
        return k1, k2, probs, (E[0] if E.shape[0] == 1 else E)
