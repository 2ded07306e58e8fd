#!/usr/bin/env python3
"""Within-process spread of the step time: 30 windows of 10 steps each, plus the host enqueue time of each window.
usage: python tools/step_variance.py [dense|sparse] [batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import onnx_image_processing_b200 as om
from oracle import oracle as O   # synthetic inputs only

wl = sys.argv[1] if len(sys.argv) > 1 else "dense"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cls = {"dense": om.ShiTomasiBADSinkhornMatcher, "sparse": om.ShiTomasiSparseBADSinkhornMatcher}[wl]
model = cls(512).cuda().eval()
i1, i2 = (t.cuda() for t in O.texture_images(B, 480, 640, seed=3))
rows = []
mode = sys.argv[3] if len(sys.argv) > 3 else "plain"     # plain | nvml (NVML initialised and primed) | pinned (big pinned buffers alive)
if mode == "nvml":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    sampler = bench.ClockSampler(0)
if mode == "pinned":
    hp = [torch.empty(B, 480, 640).pin_memory() for _ in range(2)]
with torch.no_grad():
    for _ in range(3): model(i1, i2)
    torch.cuda.synchronize()
    for w in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record()
        for _ in range(10): model(i1, i2)
        b.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
        rows.append((a.elapsed_time(b) / 10, 1e3 * (t1 - t0) / 10))
print(wl, mode, "GPU ms/step per window:", " ".join(f"{g:.3f}" for g, _ in rows))
print(wl, "host enqueue ms/step:  ", " ".join(f"{h:.3f}" for _, h in rows))
