#!/usr/bin/env python3
"""A few calls of the descriptor stages (dense integral + keypoint kernel, sparse stage) and the detector on 64 images of
480x640, for `ncu --metrics gpu__time_duration.sum` launch lists (tools/gpu_ncu_list.sh)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import onnx_image_processing_b200 as om
from onnx_image_processing_b200 import _native as nat, _ops
from oracle import oracle as O

B, H, W, K, P = 64, 480, 640, 512, 256
dev = "cuda:0"
u8 = len(sys.argv) > 1 and sys.argv[1] == "u8"
i1, _ = O.texture_images(B, H, W, seed=1000)
img = i1.to(dev)
if u8:
    img = img.to(torch.uint8)
reps = 3
with torch.no_grad():
    for _ in range(reps):
        k, s = _ops.detect(img, K, 3, 3, 0.0, 0)
    tb = om.BADDescriptor()._pair_table.to(dev)
    if not u8:
        for _ in range(reps):
            _ops.dense_bad_at_keypoints(img, k, tb, 0, 10.0, True)
        for _ in range(reps):
            om.SparseBAD().to(dev)(img, k)
    m = om.ShiTomasiBADSinkhornMatcher(K).to(dev).eval()
    for _ in range(reps):
        m(img, img)
torch.cuda.synchronize()
print("done")
