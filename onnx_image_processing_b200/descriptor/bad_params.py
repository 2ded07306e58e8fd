"""Learned BAD measurement tables (data of pytorch_model/descriptor/bad_params.py:4-1568).

The values live in data/bad_tables.npz (and, identically, in csrc/bad_tables.inc for the C side);
tools/gen_bad_tables.py regenerates both from a checkout of the reference.
"""
import os

import numpy as np
import torch

_NPZ = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "bad_tables.npz")


def _get_bad_learned_params(num_pairs: int) -> tuple[torch.Tensor, torch.Tensor]:
    """(num_pairs,5) float32 rows (x1, x2, y1, y2, radius) in 32x32 patch coordinates, and thresholds."""
    if num_pairs not in (256, 512):
        raise ValueError(f"num_pairs must be 256 or 512 to use learned BAD patterns, got {num_pairs}")
    z = np.load(_NPZ)
    box = z[f"boxes{num_pairs}"].astype(np.float32)
    box[:, :4] += 16.0
    return torch.from_numpy(box), torch.from_numpy(z[f"thresholds{num_pairs}"].astype(np.float32))
