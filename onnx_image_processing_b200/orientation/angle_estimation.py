"""AngleEstimator: drop-in for pytorch_model/orientation/angle_estimation.py:28-172."""
import torch
from torch import nn

from .. import _ops


class AngleEstimator(nn.Module):
    """Per-pixel orientation atan2(m01, m10) from Gaussian-weighted intensity moments."""

    def __init__(self, patch_size: int = 15, sigma: float = 2.5):
        super().__init__()
        if patch_size % 2 == 0:
            raise ValueError(f"patch_size must be odd, got {patch_size}")
        if sigma <= 0:
            raise ValueError(f"sigma must be positive, got {sigma}")
        self.patch_size = patch_size
        self.sigma = sigma
        half = patch_size // 2
        ax = torch.arange(-half, half + 1, dtype=torch.float32)
        y, x = torch.meshgrid(ax, ax, indexing="ij")
        g = torch.exp(-(x ** 2 + y ** 2) / (2 * sigma ** 2))          # angle_estimation.py:108
        self.register_buffer("moment_kernels", torch.stack([x * g, y * g]).unsqueeze(1))   # (2,1,ps,ps)

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        return _ops.angle_map(image, self.moment_kernels)
