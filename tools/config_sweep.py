#!/usr/bin/env python3
"""Device-resident pairs/s for the BASELINE configs that are not the bench default:
config 3 (sparse matcher, batch sweep), config 4 (rotation-invariant matcher), config 5 (1080x1920, K=2048),
and the export defaults (num_pairs 512, hard binarisation, epsilon 0.05, nms_radius 5)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import onnx_image_processing_b200 as om
from oracle import oracle as O

def run(model, i1, i2, steps):
    with torch.no_grad():
        for _ in range(3): model(i1, i2)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(steps): model(i1, i2)
        b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps

out = []
base1, base2 = O.texture_images(64, 480, 640, seed=1)
def imgs(B):
    reps = (B + 63) // 64
    return torch.cat([base1] * reps)[:B].cuda(), torch.cat([base2] * reps)[:B].cuda()
m = om.ShiTomasiSparseBADSinkhornMatcher(512).cuda().eval()
for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024):
    i1, i2 = imgs(B)
    ms = run(m, i1, i2, 20 if B <= 128 else 5)
    out.append(dict(config="3: sparse matcher 480x640 k=512", batch=B, ms=ms, pairs_per_s=B / ms * 1e3))
    del i1, i2
i1, i2 = imgs(64)
for name, mod in (("4: angle matcher 480x640 k=512", om.ShiTomasiAngleSparseBADSinkhornMatcher(512)),
                  ("2: dense matcher 480x640 k=512", om.ShiTomasiBADSinkhornMatcher(512)),
                  ("3 at export defaults (P=512, hard, eps 0.05, nms 5, k=1024)",
                   om.ShiTomasiSparseBADSinkhornMatcher(1024, num_pairs=512, binarize=True, soft_binarize=False, epsilon=0.05, nms_radius=5))):
    ms = run(mod.cuda().eval(), i1, i2, 10)
    out.append(dict(config=name, batch=64, ms=ms, pairs_per_s=64 / ms * 1e3))
del i1, i2
b1, b2 = O.texture_images(8, 1080, 1920, seed=2)
m5 = om.ShiTomasiSparseBADSinkhornMatcher(2048).cuda().eval()
for B in (1, 8):
    ms = run(m5, b1[:B].cuda(), b2[:B].cuda(), 5)
    out.append(dict(config="5: sparse matcher 1080x1920 k=2048 (generic Sinkhorn path)", batch=B, ms=ms, pairs_per_s=B / ms * 1e3))
for r in out:
    print(json.dumps(r))
