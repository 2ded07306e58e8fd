#!/usr/bin/env python3
"""Host-side cost of one matcher step: time to ENQUEUE steps (no synchronisation) vs GPU time per step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import onnx_image_processing_b200 as om
from oracle import oracle as O
wl = sys.argv[1] if len(sys.argv) > 1 else "dense"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cls = {"dense": om.ShiTomasiBADSinkhornMatcher, "sparse": om.ShiTomasiSparseBADSinkhornMatcher}[wl]
model = cls(512).cuda().eval()
i1, i2 = O.texture_images(B, 480, 640, seed=3)
i1, i2 = i1.cuda(), i2.cuda()
with torch.no_grad():
    for _ in range(5): model(i1, i2)
    torch.cuda.synchronize()
    for steps in (20, 40):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record()
        for _ in range(steps): model(i1, i2)
        b.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"{wl} B={B} steps={steps}: host enqueue {1e3 * (t1 - t0) / steps:.3f} ms/step, GPU {a.elapsed_time(b) / steps:.3f} ms/step")
