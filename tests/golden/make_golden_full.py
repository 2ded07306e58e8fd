#!/usr/bin/env python3
"""Mint the FULL-SIZE golden vectors from the LIVE reference (authoring container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_full.py [/root/reference]

  dense_full_default    BASELINE configs[1]: ShiTomasiBADSinkhornMatcher(512) defaults at 480x640 (the headline bench)
  sparse_full_export    the configuration the reference's export script ships
                        (onnx_export/export_shi_tomasi_sparse_bad_sinkhorn.py:52-127): K=1024, 512 pairs, hard
                        binarisation, epsilon 0.05, NMS radius 5 at 480x640 -> generic (K > 512) Sinkhorn, radius-5
                        detector routing and the 512-pair tables end to end
  sparse_1080p_k2048    BASELINE configs[4]: 1080x1920, K=2048, both images, descriptors and P
  sinkhorn_with_scores  SinkhornMatcherWithScores (matching/sinkhorn.py:211-259)
  matches_dense_full    MatchExtractionWrapper over the headline matcher (feature_detection/match_extraction_wrapper.py)

Large matrices are stored as a row sample (P_rows: the row indices, P_sample: those full rows) plus the complete
dustbin row / column, every row's maximum and argmax and the float64 sum; everything else is stored whole.
Nothing at test time reads /root/reference.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, sys.argv[1] if len(sys.argv) > 1 else "/root/reference")

from oracle import oracle as O  # noqa: E402  (only for the shared synthetic-input generators)

from pytorch_model.feature_detection.shi_tomasi_sparse_bad_sinkhorn import ShiTomasiSparseBADSinkhornMatcher  # noqa: E402
from pytorch_model.feature_detection.shi_tomasi_bad_sinkhorn import ShiTomasiBADSinkhornMatcher  # noqa: E402
from pytorch_model.feature_detection.match_extraction_wrapper import MatchExtractionWrapper  # noqa: E402
from pytorch_model.matching.sinkhorn import SinkhornMatcherWithScores  # noqa: E402

EXPORT = dict(num_pairs=512, binarize=True, soft_binarize=False, epsilon=0.05, nms_radius=5)  # export script defaults


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.numpy()
        out[k] = v
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB", flush=True)


def p_summary(p, n_rows, seed):
    """Row sample + dustbin row / column + per-row max / argmax + float64 sum of a (1, N+1, M+1) matrix."""
    N = p.shape[1] - 1
    g = torch.Generator().manual_seed(seed)
    rows = torch.sort(torch.randperm(N, generator=g)[:n_rows]).values
    core = p[0, :N, :N]
    return dict(P_rows=rows, P_sample=p[0, rows], P_dust_row=p[0, N], P_dust_col=p[0, :, N],
                P_row_max=core.max(dim=-1).values, P_row_argmax=core.argmax(dim=-1), P_col_argmax=core.argmax(dim=-2),
                P_sum64=float(p.double().sum()))


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    only = set(sys.argv[2:])

    if not only or "dense" in only:
        i1, i2 = O.texture_images(1, 480, 640, 1000)            # seed 1000 = pair 0 of bench.py's rank-0 batch
        m = ShiTomasiBADSinkhornMatcher(max_keypoints=512).eval()
        with torch.no_grad():
            k1, k2, p = m(i1, i2)
            _, dmap = m.detector(i1)
            d1 = torch.nn.functional.normalize(m._extract_descriptors_at_keypoints_batched(dmap, k1), p=2, dim=-1)
            del dmap
            _, dmap = m.detector(i2)
            d2 = torch.nn.functional.normalize(m._extract_descriptors_at_keypoints_batched(dmap, k2), p=2, dim=-1)
            del dmap
            wr = MatchExtractionWrapper(m, max_matches=100, match_threshold=0.0035).eval()
            mk1, mk2, ms, mv = wr(i1, i2)
        save("dense_full_default", kind="dense_full", image1=i1.to(torch.uint8), image2=i2.to(torch.uint8), K=512,
             kwargs=repr({}), kpts1=k1, kpts2=k2, desc1=d1, desc2=d2, P=p)
        save("matches_dense_full", kind="matches_full", source="dense_full_default", max_matches=100, threshold=0.0035,
             mk1=mk1, mk2=mk2, scores=ms, valid=mv)

    if not only or "export" in only:
        i1, i2 = O.texture_images(1, 480, 640, 2000)
        m = ShiTomasiSparseBADSinkhornMatcher(max_keypoints=1024, **EXPORT).eval()
        with torch.no_grad():
            k1, k2, p = m(i1, i2)
            d1 = m.descriptor(i1, k1)
            d2 = m.descriptor(i2, k2)
        # hard-binarised descriptors are {0, 1} / norm: store the bits and the norms
        save("sparse_full_export", kind="sparse_full", image1=i1.to(torch.uint8), image2=i2.to(torch.uint8), K=1024,
             kwargs=repr(dict(EXPORT)), kpts1=k1, kpts2=k2, desc1_bits=np.packbits((d1 > 0).numpy(), axis=-1),
             desc2_bits=np.packbits((d2 > 0).numpy(), axis=-1), desc1_norm=d1.max(dim=-1).values,
             desc2_norm=d2.max(dim=-1).values, **p_summary(p, 192, 1))

    if not only or "1080p" in only:
        i1, i2 = O.texture_images(1, 1080, 1920, 3000)
        m = ShiTomasiSparseBADSinkhornMatcher(max_keypoints=2048).eval()
        with torch.no_grad():
            k1, k2, p = m(i1, i2)
            d1 = m.descriptor(i1, k1)
            d2 = m.descriptor(i2, k2)
        g = torch.Generator().manual_seed(2)
        rows = torch.sort(torch.randperm(2048, generator=g)[:256]).values
        save("sparse_1080p_k2048", kind="sparse_full", image1=i1.to(torch.uint8), image2=i2.to(torch.uint8), K=2048,
             kwargs=repr({}), kpts1=k1, kpts2=k2, desc_rows=rows, desc1_sample=d1[0, rows], desc2_sample=d2[0, rows],
             desc1_sum64=float(d1.double().sum()), desc2_sum64=float(d2.double().sum()),
             desc1_rowsum=d1[0].double().sum(dim=-1).float(), desc2_rowsum=d2[0].double().sum(dim=-1).float(),
             **p_summary(p, 96, 3))

    if not only or "scores" in only:
        g = torch.Generator().manual_seed(13)
        d1 = torch.nn.functional.normalize(torch.randn(2, 200, 256, generator=g), dim=-1)
        d2 = torch.nn.functional.normalize(d1[:, torch.randperm(200, generator=g)[:160]] + 0.3 * torch.randn(2, 160, 256, generator=g), dim=-1)
        kw = dict(iterations=20, epsilon=0.1)
        with torch.no_grad():
            p, s0, s1 = SinkhornMatcherWithScores(**kw)(d1, d2)
        save("sinkhorn_with_scores", kind="with_scores", desc1=d1, desc2=d2, kwargs=repr(kw), P=p, scores0=s0, scores1=s1)


if __name__ == "__main__":
    main()
