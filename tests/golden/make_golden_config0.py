#!/usr/bin/env python3
"""Mint the golden vectors of BASELINE configs[0] from the LIVE reference (authoring container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_config0.py [/root/reference]

  config0_detector_480x640_k1000   "Shi-Tomasi + BAD detection, single 480x640 grayscale image, max_keypoints=1000,
                                   PyTorch CPU module": both forms SURVEY.md names for it, on one seeded image --
    (A) ShiTomasiAngleSparseBADDetector(max_keypoints=1000) (feature_detection/shi_tomasi_angle.py:246-356):
        keypoints, scores, descriptors at the keypoints;
    (B) ShiTomasiBADDetector() (feature_detection/shi_tomasi_bad.py:20-89): the score map whole, 8192 probes of the dense
        (1,256,H,W) descriptor map, then the library's own selection on that score map (utils/keypoint_utils.py:12-117:
        NMS radius 3, threshold 0.01, top 1000 -- the numbers of README.md:61-63) and the dense map's descriptors at those
        keypoints.
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

from oracle import oracle as O  # noqa: E402  (only for the shared synthetic-input generator)

from pytorch_model.feature_detection.shi_tomasi_angle import ShiTomasiAngleSparseBADDetector  # noqa: E402
from pytorch_model.feature_detection.shi_tomasi_bad import ShiTomasiBADDetector  # noqa: E402
from pytorch_model.utils.keypoint_utils import apply_nms_maxpool, select_topk_keypoints  # noqa: E402

K, H, W, SEED = 1000, 480, 640, 2000
NMS_RADIUS, THRESHOLD = 3, 0.01


def main():
    img = O.texture_images(1, H, W, SEED)[0]
    with torch.no_grad():
        ak, asc, ad = ShiTomasiAngleSparseBADDetector(max_keypoints=K).eval()(img)
        sc, dmap = ShiTomasiBADDetector().eval()(img)
        s3 = sc.squeeze(1)
        bk, bs = select_topk_keypoints(s3, apply_nms_maxpool(s3, NMS_RADIUS), K, THRESHOLD, 0)
        yi, xi = bk[0, :, 0].long().clamp(min=0), bk[0, :, 1].long().clamp(min=0)
        bd = dmap[0][:, yi, xi].T.contiguous()                     # (K, 256) raw dense-map values at the keypoints
        bd[bk[0, :, 0] < 0] = 0.0
    g = torch.Generator().manual_seed(SEED + 99)
    n = 8192
    pp, py, px = (torch.randint(0, hi, (n,), generator=g) for hi in (dmap.shape[1], H, W))
    py[:256] = 0; py[256:512] = H - 1; px[512:768] = 0; px[768:1024] = W - 1      # some probes on the borders
    out = dict(kind="config0", image1=img.to(torch.uint8), K=K, nms_radius=NMS_RADIUS, threshold=THRESHOLD,
               a_kpts=ak, a_scores=asc, a_desc=ad, score_map=sc, b_kpts=bk, b_scores=bs, b_desc=bd,
               probe_idx=torch.stack([pp, py, px]), probe_val=dmap[0, pp, py, px])
    path = os.path.join(HERE, "config0_detector_480x640_k1000.npz")
    np.savez_compressed(path, **{k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()})
    print(f"{path}: {os.path.getsize(path) / 1e6:.2f} MB; valid keypoints A {int((ak[..., 0] >= 0).sum())}, B {int((bk[..., 0] >= 0).sum())}")


if __name__ == "__main__":
    main()
