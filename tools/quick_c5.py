"""Device-resident step of BASELINE config 5 (1080x1920, K=2048, batch 8 and 1) and of the export defaults (batch 64)."""
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

b1, b2 = O.texture_images(8, 1080, 1920, seed=2)
m5 = om.ShiTomasiSparseBADSinkhornMatcher(2048).cuda().eval()
for B in (1, 8):
    ms = run(m5, b1[:B].cuda(), b2[:B].cuda(), 10)
    print(f"config 5 batch {B}: {ms:.3f} ms  {B / ms * 1e3:.0f} pairs/s")
if "--export" in sys.argv:
    base1, base2 = O.texture_images(64, 480, 640, seed=1)
    m = om.ShiTomasiSparseBADSinkhornMatcher(1024, num_pairs=512, binarize=True, soft_binarize=False, epsilon=0.05, nms_radius=5).cuda().eval()
    ms = run(m, base1.cuda(), base2.cuda(), 10)
    print(f"export defaults batch 64: {ms:.3f} ms  {64 / ms * 1e3:.0f} pairs/s")
