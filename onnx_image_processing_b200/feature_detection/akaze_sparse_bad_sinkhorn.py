"""AKAZESparseBADSinkhornMatcher: drop-in for pytorch_model/feature_detection/akaze_sparse_bad_sinkhorn.py:27-196.

Everything after the detector's score / orientation maps -- NMS, top-k, orientation-aware sparse BAD at the keypoints and
Sinkhorn -- is one C call (om_match_pairs_from_maps_f32) into the same kernels the Shi-Tomasi matchers use."""
import torch
from torch import nn

from .. import _ops
from ..descriptor.bad import SparseBAD
from ..detector.akaze import AKAZE
from ..matching.sinkhorn import SinkhornMatcher


class AKAZESparseBADSinkhornMatcher(nn.Module):
    """Two grayscale images (B,1,H,W) -> keypoints1, keypoints2 (B,K,2) and match probabilities (B,K+1,K+1).

    ``detector`` may be replaced by any module returning ``(scores, orientations)`` maps of shape (B,1,H,W)."""

    def __init__(self, max_keypoints: int, num_scales: int = 3, diffusion_iterations: int = 3, kappa: float = 0.05,
                 threshold: float = 0.001, akaze_nms_size: int = 5, orientation_patch_size: int = 15,
                 orientation_sigma: float = 2.5, num_pairs: int = 256, binarize: bool = False, soft_binarize: bool = True,
                 temperature: float = 10.0, sinkhorn_iterations: int = 20, epsilon: float = 1.0, unused_score: float = 1.0,
                 distance_type: str = "l2", nms_radius: int = 3, score_threshold: float = 0.0,
                 normalize_descriptors: bool = True, sampling_mode: str = "nearest", border_margin: int | None = None) -> None:
        super().__init__()
        self.max_keypoints = max_keypoints
        self.nms_radius = nms_radius
        self.score_threshold = score_threshold
        self.detector = AKAZE(num_scales=num_scales, diffusion_iterations=diffusion_iterations, kappa=kappa,
                              threshold=threshold, nms_size=akaze_nms_size, orientation_patch_size=orientation_patch_size,
                              orientation_sigma=orientation_sigma)
        self.descriptor = SparseBAD(num_pairs=num_pairs, binarize=binarize, soft_binarize=soft_binarize,
                                    temperature=temperature, normalize_descriptors=normalize_descriptors,
                                    sampling_mode=sampling_mode)
        self.border_margin = self.descriptor.max_radius if border_margin is None else border_margin   # reference :131-134
        self.matcher = SinkhornMatcher(iterations=sinkhorn_iterations, epsilon=epsilon, unused_score=unused_score,
                                       distance_type=distance_type)

    def match_from_maps(self, image1, image2, scores1, scores2, orient1, orient2):
        """(kpts1, kpts2, probs, desc1, desc2) from caller-provided detector maps (reference forward :155-196)."""
        d, m = self.descriptor, self.matcher
        return _ops.match_pairs_from_maps(
            image1, image2, scores1, scores2, orient1, orient2, d._pair_table, int(self.max_keypoints), int(self.nms_radius),
            int(self.border_margin), float(self.score_threshold), d._mode(), float(d.temperature),
            bool(d.normalize_descriptors), _ops.sampling_code(d.sampling_mode), m.iterations, float(m.epsilon),
            float(m.unused_score), m.distance_type == "l1")

    def forward(self, image1: torch.Tensor, image2: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        if not image1.is_cuda:
            raise RuntimeError("image1 must be a CUDA tensor: onnx_image_processing_b200 has no CPU path")
        image1, image2 = image1.float(), image2.float()
        scores1, orient1 = self.detector(image1)
        scores2, orient2 = self.detector(image2)
        k1, k2, probs, _, _ = self.match_from_maps(image1, image2, scores1, scores2, orient1, orient2)
        return k1, k2, probs
