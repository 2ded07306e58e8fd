"""AKAZE score / orientation maps: the detector in front of ``AKAZESparseBADSinkhornMatcher`` (stands in for
pytorch_model/detector/akaze.py:25-453; same constructor arguments, sub-module and buffer names, so state_dicts interchange).

NOT part of the accelerated path: SURVEY 8(f4) asks for the AKAZE matcher to *reuse* the keypoint-selection, descriptor and
matching kernels, which start at the score / orientation maps.  The maps themselves are produced here with stock torch
operators on the input's device (a handful of 3x3 convolutions and element-wise steps per scale) -- any module with the same
``forward(image) -> (scores, orientations)`` contract can be passed to the matcher instead.
"""
import torch
import torch.nn.functional as F
from torch import nn


def _stencil(rows, scale):
    return (torch.tensor(rows, dtype=torch.float32) / scale).view(1, 1, 3, 3)


class NonLinearDiffusion(nn.Module):
    """Perona-Malik (g2) diffusion, explicit steps of dt = 0.25 (akaze.py:25-134)."""

    def __init__(self, num_iterations: int = 3, kappa: float = 0.05):
        super().__init__()
        self.num_iterations, self.kappa, self.dt = num_iterations, kappa, 0.25
        gx = _stencil([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], 8.0)
        gy = _stencil([[-1, -2, -1], [0, 0, 0], [1, 2, 1]], 8.0)
        self.register_buffer("sobel_xy", torch.cat([gx, gy]))
        self.register_buffer("sobel_xy_grouped", torch.cat([gx, gy]))

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        level = image
        for _ in range(self.num_iterations):
            grad = F.conv2d(level, self.sobel_xy, padding=1)                              # :73-83
            magnitude = torch.sqrt((grad * grad).sum(dim=1, keepdim=True) + 1e-8)         # :116
            conduction = 1.0 / (1.0 + (magnitude / self.kappa) ** 2)                      # :85-97
            div = F.conv2d(conduction * grad, self.sobel_xy_grouped, padding=1, groups=2).sum(dim=1, keepdim=True)
            level = level + self.dt * div                                                 # :131
        return level


class HessianDetector(nn.Module):
    """det(Hessian) response, kept where it is a local maximum above the threshold (akaze.py:137-263)."""

    def __init__(self, threshold: float = 0.001, nms_size: int = 5):
        super().__init__()
        self.threshold, self.nms_size = threshold, nms_size
        self.register_buffer("hessian_kernels", torch.cat([
            _stencil([[1, -2, 1], [2, -4, 2], [1, -2, 1]], 16.0),
            _stencil([[1, 2, 1], [-2, -4, -2], [1, 2, 1]], 16.0),
            _stencil([[1, 0, -1], [0, 0, 0], [-1, 0, 1]], 4.0)]))

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        h = F.conv2d(image, self.hessian_kernels, padding=1)
        response = h[:, 0:1] * h[:, 1:2] - h[:, 2:3] * h[:, 2:3]                          # :203
        peak = F.max_pool2d(response, kernel_size=self.nms_size, stride=1, padding=self.nms_size // 2)
        keep = (response == peak).float() * (response > self.threshold).float()           # :231-255
        return torch.clamp(response * keep, min=0.0)


class OrientationEstimator(nn.Module):
    """Intensity-centroid orientation with Gaussian weights, zero padding (akaze.py:266-328)."""

    def __init__(self, patch_size: int = 15, sigma: float = 2.5):
        super().__init__()
        self.patch_size, self.sigma = patch_size, sigma
        c = torch.arange(-(patch_size // 2), patch_size // 2 + 1, dtype=torch.float32)
        y, x = torch.meshgrid(c, c, indexing="ij")
        g = torch.exp(-(x ** 2 + y ** 2) / (2 * sigma ** 2))
        self.register_buffer("moment_kernels", torch.stack([x * g, y * g]).unsqueeze(1))

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        m = F.conv2d(image, self.moment_kernels, padding=self.patch_size // 2)
        return torch.atan2(m[:, 1:2], m[:, 0:1])


class AKAZE(nn.Module):
    """(B,1,H,W) -> (scores, orientations), both (B,1,H,W): best response over the scales and the orientation of the scale(s)
    that attain it (akaze.py:331-453)."""

    def __init__(self, num_scales: int = 3, diffusion_iterations: int = 3, kappa: float = 0.05, threshold: float = 0.001,
                 nms_size: int = 5, orientation_patch_size: int = 15, orientation_sigma: float = 2.5):
        super().__init__()
        self.num_scales = num_scales
        self.diffusion_layers = nn.ModuleList(
            [NonLinearDiffusion(num_iterations=diffusion_iterations, kappa=kappa) for _ in range(num_scales)])
        self.detector = HessianDetector(threshold=threshold, nms_size=nms_size)
        self.orientation_estimator = OrientationEstimator(patch_size=orientation_patch_size, sigma=orientation_sigma)

    def forward(self, image: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        level, responses, angles = image, [], []
        for layer in self.diffusion_layers:
            level = layer(level)
            responses.append(self.detector(level))
            angles.append(self.orientation_estimator(level))
        responses, angles = torch.stack(responses), torch.stack(angles)
        scores = responses.amax(dim=0)
        pick = (responses == scores.unsqueeze(0)).float()                                 # ties share the weight, :441-449
        pick = pick / pick.sum(dim=0, keepdim=True).clamp(min=1.0)
        return scores, (angles * pick).sum(dim=0)
