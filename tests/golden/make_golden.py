#!/usr/bin/env python3
"""Mint golden vectors from the LIVE reference (authoring container only).

The reference ships no known-answer tests for this path, so parity is pinned on outputs of the
reference's own modules, run here on CPU:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py
Each case stores its uint8 input images, the constructor kwargs and every output tensor in
tests/golden/<case>.npz.  Nothing at test time reads /root/reference.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, sys.argv[1] if len(sys.argv) > 1 else "/root/reference")

from oracle import oracle as O  # noqa: E402  (only for the shared synthetic-input generators)

from pytorch_model.detector.shi_tomasi import ShiTomasiScore  # noqa: E402
from pytorch_model.utils.keypoint_utils import apply_nms_maxpool, select_topk_keypoints  # noqa: E402
from pytorch_model.feature_detection.shi_tomasi_sparse_bad_sinkhorn import ShiTomasiSparseBADSinkhornMatcher  # noqa: E402
from pytorch_model.feature_detection.shi_tomasi_bad_sinkhorn import ShiTomasiBADSinkhornMatcher  # noqa: E402
from pytorch_model.feature_detection.shi_tomasi_bad import ShiTomasiBADDetector  # noqa: E402
from pytorch_model.feature_detection.shi_tomasi_angle import ShiTomasiAngleSparseBADDetector  # noqa: E402
from pytorch_model.feature_detection.shi_tomasi_angle_sparse_bad_sinkhorn import ShiTomasiAngleSparseBADSinkhornMatcher  # noqa: E402
from pytorch_model.orientation.angle_estimation import AngleEstimator  # noqa: E402
from pytorch_model.matching.sinkhorn import SinkhornMatcher  # noqa: E402

EXPORT = dict(num_pairs=512, binarize=True, soft_binarize=False, epsilon=0.05, nms_radius=5)  # export script defaults


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def rotated(img, degrees):
    """image2 for the rotation-invariant cases: rotate about the centre, keep integer grey levels."""
    B, _, H, W = img.shape
    a = np.deg2rad(degrees)
    theta = torch.tensor([[np.cos(a), -np.sin(a) * H / W, 0.0], [np.sin(a) * W / H, np.cos(a), 0.0]],
                         dtype=torch.float32).unsqueeze(0).expand(B, -1, -1)
    grid = torch.nn.functional.affine_grid(theta, img.shape, align_corners=False)
    return torch.round(torch.nn.functional.grid_sample(img, grid, mode="bilinear", padding_mode="border",
                                                        align_corners=False)).clamp(0, 255)


def stage_outputs(model_detector, img, nms_radius, K, thr, margin):
    """Per-stage intermediates of one image, straight from the reference's stage functions."""
    sc = model_detector(img).squeeze(1)
    m = apply_nms_maxpool(sc, nms_radius)
    kp, ks = select_topk_keypoints(sc, m, K, thr, margin)
    return sc, m, kp, ks


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.numpy()
        out[k] = v
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")


def sparse_case(name, H, W, K, B, seed, kwargs, family="T"):
    if family == "T":
        i1, i2 = O.texture_images(B, H, W, seed)
    else:
        i1 = O.noise_images(B, H, W, seed)
        i2 = torch.roll(i1, (2, -3), (2, 3))
    m = ShiTomasiSparseBADSinkhornMatcher(max_keypoints=K, **kwargs).eval()
    with torch.no_grad():
        k1, k2, p = m(i1, i2)
        sc, mask, kp, ks = stage_outputs(m.corner_detector, i1, m.nms_radius, K, m.score_threshold, m.border_margin)
        d1 = m.descriptor(i1, k1)
        d2 = m.descriptor(i2, k2)
    assert torch.equal(kp, k1)
    save(name, kind="sparse", image1=i1.to(torch.uint8), image2=i2.to(torch.uint8), K=K, kwargs=repr(kwargs),
         kpts1=k1, kpts2=k2, kpt_scores1=ks, desc1=d1, desc2=d2, P=p,
         score_sha=sha(sc), mask_sha=sha(mask), score_sum=float(sc.double().sum()),
         n_candidates=int((sc * mask > 0).sum()))


def angle_case(name, H, W, K, B, seed, kwargs, deg=17.0):
    i1, _ = O.texture_images(B, H, W, seed)
    i2 = rotated(i1, deg)
    m = ShiTomasiAngleSparseBADSinkhornMatcher(max_keypoints=K, **kwargs).eval()
    det = ShiTomasiAngleSparseBADDetector(max_keypoints=K, **{k: v for k, v in kwargs.items()
                                                                 if k not in ("epsilon", "sinkhorn_iterations")}).eval()
    with torch.no_grad():
        k1, k2, p = m(i1, i2)
        _, a1 = m.detector(i1)
        _, a2 = m.detector(i2)
        d1 = m.descriptor(i1, k1, a1)
        d2 = m.descriptor(i2, k2, a2)
        dk, dsc, dd = det(i1)
    # orientation at the keypoints (what the descriptor consumes, bad.py:493-499)
    yi = k1[..., 0].clamp(min=0).long()
    xi = k1[..., 1].clamp(min=0).long()
    th1 = a1[:, 0][torch.arange(B)[:, None], yi, xi]
    save(name, kind="angle", image1=i1.to(torch.uint8), image2=i2.to(torch.uint8), K=K, kwargs=repr(kwargs),
         kpts1=k1, kpts2=k2, desc1=d1, desc2=d2, P=p, theta1_at_kpts=th1, angle_sha=sha(a1),
         det_kpts=dk, det_scores=dsc, det_desc=dd)


def dense_case(name, H, W, K, B, seed, kwargs):
    i1, i2 = O.texture_images(B, H, W, seed)
    m = ShiTomasiBADSinkhornMatcher(max_keypoints=K, **kwargs).eval()
    det = ShiTomasiBADDetector(**{k: v for k, v in kwargs.items()
                                  if k in ("block_size", "num_pairs", "binarize", "soft_binarize", "temperature")}).eval()
    with torch.no_grad():
        k1, k2, p = m(i1, i2)
        sc, dmap = det(i1)
        d1 = m._extract_descriptors_at_keypoints_batched(dmap, k1)
        if m.normalize_descriptors:
            d1 = torch.nn.functional.normalize(d1, p=2, dim=-1)
    g = torch.Generator().manual_seed(seed + 99)
    npts = 4096
    pb = torch.randint(0, B, (npts,), generator=g)
    pp = torch.randint(0, dmap.shape[1], (npts,), generator=g)
    py = torch.randint(0, H, (npts,), generator=g)
    px = torch.randint(0, W, (npts,), generator=g)
    # force some probes onto the borders
    py[:256] = 0; py[256:512] = H - 1; px[512:768] = 0; px[768:1024] = W - 1
    save(name, kind="dense", image1=i1.to(torch.uint8), image2=i2.to(torch.uint8), K=K, kwargs=repr(kwargs),
         kpts1=k1, kpts2=k2, desc1=d1, P=p, score_sha=sha(sc), dmap_sha=sha(dmap),
         probe_idx=torch.stack([pb, pp, py, px]), probe_val=dmap[pb, pp, py, px])


def sinkhorn_case(name, N, M, D, B, seed, kwargs):
    g = torch.Generator().manual_seed(seed)
    d1 = torch.nn.functional.normalize(torch.randn(B, N, D, generator=g), dim=-1)
    pick = (torch.randperm(max(N, M), generator=g) % N)[:M]
    d2 = torch.nn.functional.normalize(d1[:, pick] + 0.3 * torch.randn(B, M, D, generator=g), dim=-1)
    d1[:, -3:] = 0.0  # invalid (zero) descriptors participate unmasked, sinkhorn.py has no mask
    with torch.no_grad():
        p = SinkhornMatcher(**kwargs)(d1, d2)
        p64 = SinkhornMatcher(**kwargs)(d1.double(), d2.double())
    save(name, kind="sinkhorn", desc1=d1, desc2=d2, kwargs=repr(kwargs), P=p, P64=p64)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    sparse_case("sparse_small_default", 96, 128, 64, 2, 1, {})
    sparse_case("sparse_small_export", 96, 128, 64, 2, 2, dict(EXPORT))
    sparse_case("sparse_small_noise_bilinear", 72, 104, 48, 1, 3, dict(sampling_mode="bilinear", binarize=True), family="N")
    sparse_case("sparse_small_ragged", 67, 93, 200, 1, 4, dict(block_size=5, nms_radius=2, score_threshold=5000.0,
                                                               border_margin=3, normalize_descriptors=False))
    sparse_case("sparse_full_default", 480, 640, 512, 1, 5, {})
    angle_case("angle_small_default", 96, 128, 64, 2, 6, {})
    angle_case("angle_full_default", 480, 640, 512, 1, 7, {})
    dense_case("dense_small_default", 64, 96, 48, 1, 8, {})
    dense_case("dense_small_soft", 64, 96, 48, 1, 9, dict(binarize=True, num_pairs=512, epsilon=0.05))
    sinkhorn_case("sinkhorn_eps1", 128, 128, 256, 2, 10, dict(iterations=20, epsilon=1.0))
    sinkhorn_case("sinkhorn_eps005", 128, 96, 256, 1, 11, dict(iterations=20, epsilon=0.05))
    sinkhorn_case("sinkhorn_l1", 40, 56, 64, 1, 12, dict(iterations=7, epsilon=2.0, unused_score=0.5, distance_type="l1"))
    # constant image: zero valid keypoints -> all (-1,-1), zero descriptors
    i = torch.full((1, 1, 64, 80), 77.0)
    m = ShiTomasiSparseBADSinkhornMatcher(max_keypoints=32).eval()
    with torch.no_grad():
        k1, k2, p = m(i, i)
    save("sparse_constant_image", kind="sparse_const", image1=i.to(torch.uint8), K=32, kpts1=k1, kpts2=k2, P=p)


if __name__ == "__main__":
    main()
