"""Camera-frame ingest on the GPU: what sample/visual_odometry.py:65-92 (``load_image_from_array``) does on the host with
OpenCV -- BGR -> grey, bilinear resize to the model size, float32 -- as one kernel (om_preprocess_u8) with OpenCV's integer
arithmetic, so that the matcher sees the reference pipeline's pixels."""
import torch
from torch import nn

from . import _ops


def load_image_from_array(image: torch.Tensor, height: int, width: int, as_float: bool = True) -> torch.Tensor:
    """uint8 CUDA frame(s) (H,W,3) BGR / (H,W) grey, optionally batched (B,...) -> (B,1,height,width) with values in [0,255];
    float32 as the reference returns, or uint8 (``as_float=False``) for the matchers' native 8-bit ingest."""
    if image.dim() == 2 or (image.dim() == 3 and image.shape[-1] in (1, 3)):
        image = image.unsqueeze(0)
    return _ops.preprocess_u8(image, int(height), int(width), bool(as_float))


class FrameIngest(nn.Module):
    """``FrameIngest(height, width)(frames_u8)`` in front of any matcher module: frames (B,Hin,Win,3) BGR uint8 on the GPU ->
    (B,1,height,width) uint8, which the fused matchers read natively (4x fewer bytes than float32 pixels)."""

    def __init__(self, height: int, width: int, as_float: bool = False) -> None:
        super().__init__()
        self.height, self.width, self.as_float = int(height), int(width), bool(as_float)

    def forward(self, frames: torch.Tensor) -> torch.Tensor:
        return load_image_from_array(frames, self.height, self.width, self.as_float)
