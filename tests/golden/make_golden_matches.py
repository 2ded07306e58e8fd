#!/usr/bin/env python3
"""Golden vectors for match extraction, minted from the live reference (authoring container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_matches.py

Runs pytorch_model.matching.match_extraction.MutualNearestNeighborMatcher on the keypoints / probabilities of the
already committed matcher goldens and stores its outputs in tests/golden/matches_*.npz.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("OM_REFERENCE", "/root/reference"))

from pytorch_model.matching.match_extraction import MutualNearestNeighborMatcher  # noqa: E402

from tests import golden_util as G  # noqa: E402

# thresholds sit inside each golden's range of row maxima so that the threshold test, the mutual test and the
# fixed-size padding (max_matches > N for the 64-keypoint goldens) are all exercised
for src, max_matches, thr in (("sparse_full_default", 100, 0.0035), ("sparse_small_export", 100, 0.1),
                              ("sparse_small_default", 200, 0.03), ("angle_small_default", 32, 0.02),
                              ("dense_small_soft", 48, 0.9), ("sparse_small_ragged", 300, 0.004)):
    g = G.load(src)
    with torch.no_grad():
        mk1, mk2, sc, valid = MutualNearestNeighborMatcher(max_matches, thr)(g["P"], g["kpts1"], g["kpts2"])
    out = os.path.join(HERE, f"matches_{src}.npz")
    np.savez_compressed(out, kind="matches", source=src, max_matches=max_matches, threshold=thr, mk1=mk1.numpy(), mk2=mk2.numpy(),
                        scores=sc.numpy(), valid=valid.numpy())
    print(out, int(valid.sum()), "valid of", valid.numel())

# ---- outlier filters (SinkhornMatcherWithFilters): thresholds at the medians so that about half of the rows pass
from pytorch_model.matching.sinkhorn import SinkhornMatcherWithFilters  # noqa: E402

for src in G.names("sinkhorn"):
    g = G.load(src)
    N, M = g["desc1"].shape[1], g["desc2"].shape[1]
    core = g["P"][:, :N, :M]
    top2 = torch.topk(core, 2, dim=2).values
    ratio = float((top2[..., 0] / (top2[..., 1] + 1e-8)).median())
    margin = float((top2[..., 0] - g["P"][:, :N, M]).median())
    with torch.no_grad():
        pf, valid = SinkhornMatcherWithFilters(**g["kwargs"], ratio_threshold=ratio, dustbin_margin=margin)(g["desc1"], g["desc2"])
    out = os.path.join(HERE, f"filters_{src}.npz")
    np.savez_compressed(out, kind="filters", source=src, ratio_threshold=ratio, dustbin_margin=margin, valid=valid.numpy(),
                        row_sums=pf.sum(dim=2).numpy(), dust_col=pf[:, :N, M].numpy())
    print(out, int(valid.sum()), "valid of", valid.numel())
