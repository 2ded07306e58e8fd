"""Drop-ins for pytorch_model/utils/keypoint_utils.py:12-117 of the reference."""
import torch

from .. import _ops


def apply_nms_maxpool(scores: torch.Tensor, nms_radius: int) -> torch.Tensor:
    """(B,H,W) scores -> (B,H,W) 0/1 mask of (2r+1)^2 local maxima (keypoint_utils.py:12-44)."""
    return _ops.nms_mask(scores, int(nms_radius))


def select_topk_keypoints(
    scores: torch.Tensor,
    nms_mask: torch.Tensor,
    max_keypoints: int,
    score_threshold: float = 0.0,
    border_margin: int = 0,
) -> tuple[torch.Tensor, torch.Tensor]:
    """Top-k of scores*mask[*border] above the threshold (keypoint_utils.py:47-117).

    Returns keypoints (B,K,2) in (y,x) with (-1,-1) padding and their scores (B,K), sorted by
    descending score; equal scores are ordered by ascending flat index.
    """
    return _ops.select_topk(scores, nms_mask, int(max_keypoints), float(score_threshold), int(border_margin))
