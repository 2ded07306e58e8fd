"""ShiTomasiBADSinkhornMatcher: drop-in for
pytorch_model/feature_detection/shi_tomasi_bad_sinkhorn.py:23-219 (the dense-BAD matcher)."""
import torch
from torch import nn

from .. import _ops
from ..matching.sinkhorn import SinkhornMatcher
from .shi_tomasi_bad import ShiTomasiBADDetector


class ShiTomasiBADSinkhornMatcher(nn.Module):
    """Same outputs as the reference's dense matcher.  The reference materialises the (B,P,H,W)
    dense BAD map (315 MB per 480x640 image) and bilinearly samples K points of it; here the dense
    formula (float32 integral image, same tap order) is evaluated only at the four pixels around
    each keypoint, which yields the same K descriptors without the map."""

    def __init__(self, max_keypoints: int, block_size: int = 3, sobel_size: int = 3, num_pairs: int = 256,
                 binarize: bool = False, soft_binarize: bool = True, temperature: float = 10.0,
                 sinkhorn_iterations: int = 20, epsilon: float = 1.0, unused_score: float = 1.0,
                 distance_type: str = "l2", nms_radius: int = 3, score_threshold: float = 0.0,
                 normalize_descriptors: bool = True) -> None:
        super().__init__()
        self.max_keypoints = max_keypoints
        self.nms_radius = nms_radius
        self.score_threshold = score_threshold
        self.normalize_descriptors = normalize_descriptors
        self.detector = ShiTomasiBADDetector(block_size=block_size, sobel_size=sobel_size, num_pairs=num_pairs,
                                             binarize=binarize, soft_binarize=soft_binarize, temperature=temperature)
        self.matcher = SinkhornMatcher(iterations=sinkhorn_iterations, epsilon=epsilon, unused_score=unused_score,
                                       distance_type=distance_type)

    def _extract_descriptors_at_keypoints_batched(self, descriptor_map: torch.Tensor,
                                                  keypoints: torch.Tensor) -> torch.Tensor:
        """Bilinear gather from a caller-provided dense map with invalid rows zeroed (reference :120-160)."""
        _, _, H, W = descriptor_map.shape
        valid = (keypoints[:, :, 0] >= 0).to(descriptor_map.dtype)
        kc = torch.stack([keypoints[:, :, 0].clamp(0.0, float(H - 1)), keypoints[:, :, 1].clamp(0.0, float(W - 1))],
                         dim=-1)
        return _ops.gather_descriptors(descriptor_map, kc, True) * valid.unsqueeze(-1)

    def _match_args(self):
        d, m = self.detector.descriptor, self.matcher
        return (d._pair_table, None, _ops.MATCH_DENSE, int(self.max_keypoints),
                self.detector.corner_detector.block_size, int(self.nms_radius), 0,
                float(self.score_threshold), _ops.desc_mode(d.binarize, d.soft_binarize),
                float(d.temperature), bool(self.normalize_descriptors), _ops.SAMPLE_NEAREST,
                m.iterations, float(m.epsilon), float(m.unused_score), m.distance_type == "l1")

    def match(self, image1: torch.Tensor, image2: torch.Tensor):
        return _ops.match_pairs(image1, image2, *self._match_args())

    def forward(self, image1: torch.Tensor, image2: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        k1, k2, probs, _, _ = self.match(image1, image2)
        return k1, k2, probs
