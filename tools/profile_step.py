#!/usr/bin/env python3
"""Two steps of one matcher over a synthetic batch -- the command line profiled by ncu.
usage: python tools/profile_step.py [dense|sparse|angle] [batch] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import onnx_image_processing_b200 as om  # noqa: E402
from oracle import oracle as O  # noqa: E402  (synthetic inputs only)

wl = sys.argv[1] if len(sys.argv) > 1 else "dense"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cls = {"dense": om.ShiTomasiBADSinkhornMatcher, "sparse": om.ShiTomasiSparseBADSinkhornMatcher,
       "angle": om.ShiTomasiAngleSparseBADSinkhornMatcher}[wl]
model = cls(512).cuda().eval()
i1, i2 = O.texture_images(B, 480, 640, seed=3)
i1, i2 = i1.cuda(), i2.cuda()
with torch.no_grad():
    for _ in range(steps):
        k1, k2, p = model(i1, i2)
torch.cuda.synchronize()
print("ok", float(p[0, :512, :512].sum()))
