#!/usr/bin/env python3
"""Golden vectors for the essential-matrix head, minted from the live reference (authoring container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_essential.py

grid cases:   pytorch_model.geometry.essential_matrix_estimator.EssentialMatrixEstimator on seeded random P
module cases: pytorch_model.feature_detection.shi_tomasi_angle_sparse_bad_sinkhorn_essential_matrix.
              ShiTomasiAngleSparseBADSinkhornWithEssentialMatrix on texture image pairs (image 2 = image 1 shifted)
Stored in tests/golden/essential_*.npz; nothing at test time reads /root/reference.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("OM_REFERENCE", "/root/reference"))

from pytorch_model.geometry.essential_matrix_estimator import EssentialMatrixEstimator  # noqa: E402
from pytorch_model.feature_detection.shi_tomasi_angle_sparse_bad_sinkhorn_essential_matrix import (  # noqa: E402
    ShiTomasiAngleSparseBADSinkhornWithEssentialMatrix)

from oracle import oracle as O  # noqa: E402  (only for the shared synthetic-input generator)

K_GRID = torch.tensor([[16.0, 0.0, 16.0], [0.0, 16.0, 16.0], [0.0, 0.0, 1.0]])       # the reference's own demo intrinsics
K_CAM = torch.tensor([[525.0, 0.0, 320.0], [0.0, 525.0, 240.0], [0.0, 0.0, 1.0]])    # the reference's docstring example

for name, n, m, seed, kw in (("grid_64", 64, 64, 1, {}), ("grid_200x150", 200, 150, 2, {}),
                             ("grid_128x256_top2", 128, 256, 3, dict(top_k=2, n_iter=12, n_iter_manifold=4)),
                             ("grid_96_top5", 96, 96, 4, dict(top_k=5))):
    torch.manual_seed(seed)
    P = torch.rand(n + 1, m + 1)
    if "top5" in name:                                  # exact duplicates inside rows / columns: top-k counts them
        P = torch.round(P * 50) / 50
    with torch.no_grad():
        E = EssentialMatrixEstimator(K_GRID, image_shape=(32, 32), **kw)(P)
    out = os.path.join(HERE, f"essential_{name}.npz")
    np.savez_compressed(out, kind="essential_grid", K=K_GRID.numpy(), image_shape=np.array([32, 32]), kwargs=repr(kw),
                        P=P.numpy(), E=E.numpy())
    print(out, E.abs().max().item())

for name, H, W, k, seed, kw in (("module_240x320_k128", 240, 320, 128, 5, {}),
                                ("module_120x160_k64", 120, 160, 64, 7, {}),
                                ("module_96x128_k200_ragged", 96, 128, 200, 8, dict(epsilon=0.5, nms_radius=5)),
                                ("module_240x320_k256_eps02", 240, 320, 256, 9, dict(epsilon=0.2, top_k=2))):
    i1, i2 = O.texture_images(1, H, W, seed=seed)
    i1, i2 = i1.reshape(1, 1, H, W), i2.reshape(1, 1, H, W)
    with torch.no_grad():
        k1, k2, P, E = ShiTomasiAngleSparseBADSinkhornWithEssentialMatrix(K_CAM, k, **kw).eval()(i1, i2)
    out = os.path.join(HERE, f"essential_{name}.npz")
    np.savez_compressed(out, kind="essential_module", K=K_CAM.numpy(), max_keypoints=k, kwargs=repr(kw),
                        image1=i1.numpy().astype(np.uint8), image2=i2.numpy().astype(np.uint8), kpts1=k1.numpy(),
                        kpts2=k2.numpy(), P=P.numpy(), E=E.numpy())
    print(out, "valid", int((k1[0, :, 0] >= 0).sum()), int((k2[0, :, 0] >= 0).sum()), "max|E|", E.abs().max().item())
