#!/usr/bin/env python3
"""Sinkhorn kernel time vs number of pairs: shows how many 8-CTA clusters the GPU runs concurrently."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import onnx_image_processing_b200 as om
g = torch.Generator().manual_seed(0)
m = om.SinkhornMatcher(20, 1.0).cuda()
for B in (1, 8, 12, 14, 16, 17, 18, 24, 32, 48, 64, 128):
    d1 = torch.nn.functional.normalize(torch.randn(B, 512, 256, generator=g), dim=-1).cuda()
    d2 = torch.nn.functional.normalize(torch.randn(B, 512, 256, generator=g), dim=-1).cuda()
    for _ in range(3): m(d1, d2)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(10): m(d1, d2)
    b.record(); torch.cuda.synchronize()
    print(f"B={B:4d}: {a.elapsed_time(b) / 10 * 1000:8.1f} us  ({a.elapsed_time(b) / 10 * 1000 / B:6.2f} us/pair)")
