"""ShiTomasiScore: drop-in for pytorch_model/detector/shi_tomasi.py:6-112 of the reference."""
import torch
from torch import nn

from .. import _ops


class ShiTomasiScore(nn.Module):
    """Minimum-eigenvalue corner response for every pixel, (N,1,H,W) -> (N,1,H,W) float32.

    Same constructor, buffers and errors as the reference (shi_tomasi.py:34-64); forward runs the
    fused sm_100a stencil kernel (csrc/detect.cu) instead of pad + conv2d x2 + elementwise ops.
    """

    def __init__(self, block_size: int = 3, sobel_size: int = 3) -> None:
        super().__init__()
        if sobel_size != 3:
            raise ValueError(f"sobel_size must be 3, got {sobel_size}")
        if block_size <= 0 or block_size % 2 == 0:
            raise ValueError(f"block_size must be a positive odd integer, got {block_size}")
        self.block_size = block_size
        self.sobel_size = sobel_size
        # constant buffers kept for state_dict compatibility; the kernel has them baked in
        sx = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]])
        self.register_buffer("sobel_xy", torch.stack([sx, sx.t()]).unsqueeze(1))
        self.register_buffer("sum_kernel_grouped", torch.ones(3, 1, block_size, block_size))

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        return _ops.shi_tomasi_score(image, self.block_size)
