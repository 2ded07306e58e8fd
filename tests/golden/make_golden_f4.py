#!/usr/bin/env python3
"""Mint the goldens of SURVEY 8(f4) from the LIVE reference / OpenCV (authoring container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_f4.py [/root/reference]

  akaze_*   pytorch_model/feature_detection/akaze_sparse_bad_sinkhorn.py on a seeded pair: the detector's score and orientation
            maps (what the accelerated part starts from), keypoints, descriptors and P
  ingest_*  sample/visual_odometry.py:65-92 load_image_from_array (cv2.cvtColor + cv2.resize + float32) on seeded BGR frames
Nothing at test time reads /root/reference or imports cv2.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, sys.argv[1] if len(sys.argv) > 1 else "/root/reference")

import cv2  # noqa: E402

from oracle import oracle as O  # noqa: E402  (synthetic inputs)
from pytorch_model.feature_detection.akaze_sparse_bad_sinkhorn import AKAZESparseBADSinkhornMatcher  # noqa: E402
from pytorch_model.utils import apply_nms_maxpool, select_topk_keypoints  # noqa: E402


def save(name, **arrs):
    out = {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in arrs.items()}
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB", flush=True)


def akaze(name, H, W, K, seed, **kw):
    i1, i2 = O.texture_images(1, H, W, seed)
    m = AKAZESparseBADSinkhornMatcher(max_keypoints=K, **kw).eval()
    with torch.no_grad():
        k1, k2, p = m(i1, i2)
        s1, o1 = m.detector(i1)
        s2, o2 = m.detector(i2)
        d1 = m.descriptor(i1, k1, o1)
        d2 = m.descriptor(i2, k2, o2)
        # the same selection the module ran, for the keypoint scores
        _, ks1 = select_topk_keypoints(s1.squeeze(1), apply_nms_maxpool(s1.squeeze(1), m.nms_radius), K, m.score_threshold,
                                       m.border_margin)
    save(name, kind="akaze", image1=i1.to(torch.uint8), image2=i2.to(torch.uint8), K=K, kwargs=repr(kw), scores1=s1, scores2=s2,
         orient1=o1, orient2=o2, kpts1=k1, kpts2=k2, kscores1=ks1, desc1=d1, desc2=d2, P=p)


def ingest(name, Hin, Win, H, W, seed, channels=3):
    rng = np.random.default_rng(seed)
    # smooth content + noise, so that bilinear weights matter
    base = rng.integers(0, 256, (Hin // 8 + 2, Win // 8 + 2, channels), dtype=np.uint8)
    frame = cv2.resize(base, (Win, Hin), interpolation=cv2.INTER_CUBIC)
    frame = np.clip(frame.astype(np.int32) + rng.integers(-20, 21, frame.shape if channels == 3 else (Hin, Win)), 0, 255).astype(np.uint8)
    if channels == 1:
        frame = frame.reshape(Hin, Win)
    gray = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY) if channels == 3 else frame
    out = cv2.resize(gray, (W, H), interpolation=cv2.INTER_LINEAR).astype(np.float32)[np.newaxis, np.newaxis]
    save(name, kind="ingest", frame=frame, height=H, width=W, out=out, cv2_version=cv2.__version__)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    akaze("akaze_small_default", 120, 160, 96, 5)
    akaze("akaze_240x320_k256_export", 240, 320, 256, 6, num_pairs=512, binarize=True, soft_binarize=False, epsilon=0.05,
          nms_radius=5)
    ingest("ingest_720p_to_480x640", 720, 1280, 480, 640, 1)
    ingest("ingest_376x1241_to_480x640", 376, 1241, 480, 640, 3)          # KITTI-sized frame: shrinks in x, grows in y
    ingest("ingest_gray_600x800_to_481x643", 600, 800, 481, 643, 4, channels=1)
    ingest("ingest_same_size_480x640", 480, 640, 480, 640, 5)
    ingest("ingest_enlarge_300x400_to_480x640", 300, 400, 480, 640, 6)


if __name__ == "__main__":
    main()
