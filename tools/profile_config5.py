#!/usr/bin/env python3
"""Two steps of BASELINE config 5 (sparse matcher, 1080x1920, K=2048, batch 8) -- the command line profiled by ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import onnx_image_processing_b200 as om
from oracle import oracle as O   # synthetic inputs only
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
i1, i2 = (t.cuda() for t in O.texture_images(B, 1080, 1920, seed=2))
m = om.ShiTomasiSparseBADSinkhornMatcher(2048).cuda().eval()
with torch.no_grad():
    for _ in range(2):
        k1, k2, p = m(i1, i2)
torch.cuda.synchronize()
print("ok", float(p[0, :16, :16].sum()))
