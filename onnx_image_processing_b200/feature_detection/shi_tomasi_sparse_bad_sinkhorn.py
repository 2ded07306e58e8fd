"""ShiTomasiSparseBADSinkhornMatcher: drop-in for
pytorch_model/feature_detection/shi_tomasi_sparse_bad_sinkhorn.py:27-182."""
import torch
from torch import nn

from .. import _ops
from ..descriptor.bad import SparseBAD
from ..detector.shi_tomasi import ShiTomasiScore
from ..matching.sinkhorn import SinkhornMatcher


class ShiTomasiSparseBADSinkhornMatcher(nn.Module):
    """Two grayscale images (B,1,H,W) -> keypoints1, keypoints2 (B,K,2) and match probabilities (B,K+1,K+1).

    forward issues one C call (om_match_pairs_f32): fused stencil + NMS + candidate compaction,
    radix-select top-k, sparse BAD at the keypoints, similarity GEMM + cluster-resident Sinkhorn.
    """

    def __init__(self, max_keypoints: int, block_size: int = 3, sobel_size: int = 3, num_pairs: int = 256,
                 binarize: bool = False, soft_binarize: bool = True, temperature: float = 10.0,
                 sinkhorn_iterations: int = 20, epsilon: float = 1.0, unused_score: float = 1.0,
                 distance_type: str = "l2", nms_radius: int = 3, score_threshold: float = 0.0,
                 normalize_descriptors: bool = True, sampling_mode: str = "nearest",
                 border_margin: int | None = None) -> None:
        super().__init__()
        self.max_keypoints = max_keypoints
        self.nms_radius = nms_radius
        self.score_threshold = score_threshold
        self.corner_detector = ShiTomasiScore(block_size=block_size, sobel_size=sobel_size)
        self.descriptor = SparseBAD(num_pairs=num_pairs, binarize=binarize, soft_binarize=soft_binarize,
                                    temperature=temperature, normalize_descriptors=normalize_descriptors,
                                    sampling_mode=sampling_mode)
        # None -> the descriptor's largest box radius (reference :121-124)
        self.border_margin = self.descriptor.max_radius if border_margin is None else border_margin
        self.matcher = SinkhornMatcher(iterations=sinkhorn_iterations, epsilon=epsilon, unused_score=unused_score,
                                       distance_type=distance_type)

    def _match_args(self):
        d, m = self.descriptor, self.matcher
        return (d._pair_table, None, _ops.MATCH_SPARSE, int(self.max_keypoints),
                self.corner_detector.block_size, int(self.nms_radius), int(self.border_margin),
                float(self.score_threshold), d._mode(), float(d.temperature),
                bool(d.normalize_descriptors), _ops.sampling_code(d.sampling_mode), m.iterations,
                float(m.epsilon), float(m.unused_score), m.distance_type == "l1")

    def match(self, image1: torch.Tensor, image2: torch.Tensor):
        """forward plus the two descriptor sets: (kpts1, kpts2, probs, desc1, desc2)."""
        return _ops.match_pairs(image1, image2, *self._match_args())

    def forward(self, image1: torch.Tensor, image2: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        k1, k2, probs, _, _ = self.match(image1, image2)
        return k1, k2, probs
