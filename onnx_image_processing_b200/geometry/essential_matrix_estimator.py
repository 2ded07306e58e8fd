"""EssentialMatrixEstimator: drop-in for pytorch_model/geometry/essential_matrix_estimator.py:29-392."""
import torch
from torch import nn

from .. import _ops


class EssentialMatrixEstimator(nn.Module):
    """Weighted 8-point essential matrix from a Sinkhorn probability matrix whose rows / columns index the points of
    an (H, W) grid (feature i at pixel (i % W, i // W)).  Same constructor, buffers (K, K_inv, pixel_coords,
    pixel_coords_n) and `forward(P)` as the reference; `forward` additionally accepts a batch (B, N+1, M+1) and then
    returns (B, 3, 3) -- the reference handles one matrix per call.  The whole estimate is one kernel launch
    (om_essential_matrix_f32)."""

    def __init__(self, K: torch.Tensor, image_shape: tuple[int, int] = (32, 32), top_k: int = 3, n_iter: int = 30,
                 n_iter_manifold: int = 10) -> None:
        super().__init__()
        K_f = K.float()
        K_inv = torch.linalg.inv(K_f)                                         # essential_matrix_estimator.py:74-77
        self.register_buffer("K", K_f)
        self.register_buffer("K_inv", K_inv)
        self.top_k = top_k
        self.n_iter = n_iter
        self.n_iter_manifold = n_iter_manifold
        H, W = image_shape
        self.H = H
        self.W = W
        idx = torch.arange(H * W, dtype=torch.float32)                        # :88-92
        pixel_coords = torch.stack([idx % W, idx // W], dim=-1)
        self.register_buffer("pixel_coords", pixel_coords)
        pixel_coords_h = torch.cat([pixel_coords, torch.ones(H * W, 1)], dim=-1)
        self.register_buffer("pixel_coords_n", (pixel_coords_h @ K_inv.T)[:, :2])   # :101-105

    def estimate(self, P: torch.Tensor, pts1_n: torch.Tensor, pts2_n: torch.Tensor, valid1=None, valid2=None) -> torch.Tensor:
        """(B, N+1, M+1) probabilities and normalised (x, y) points -> (B, 3, 3)."""
        return _ops.essential_matrix(P, pts1_n, pts2_n, valid1, valid2, int(self.top_k), int(self.n_iter),
                                     int(self.n_iter_manifold))

    def forward(self, P: torch.Tensor) -> torch.Tensor:
        single = P.dim() == 2
        Pb = P.unsqueeze(0) if single else P
        N, M = Pb.shape[1] - 1, Pb.shape[2] - 1
        pc = self.pixel_coords_n.to(Pb.device)
        if max(N, M) > pc.shape[0]:
            raise RuntimeError(f"image_shape {self.H}x{self.W} holds {pc.shape[0]} grid points, P needs {max(N, M)}")
        E = self.estimate(Pb, pc[:N].contiguous(), pc[:M].contiguous())
        return E[0] if single else E
