#!/usr/bin/env python3
"""End-to-end pairs/s of HostBatchMatcher for a few (chunk, n_streams) settings, float32 and uint8 host images."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import onnx_image_processing_b200 as om
from onnx_image_processing_b200.host_pipeline import HostBatchMatcher
from oracle import oracle as O
B = 64
model = om.ShiTomasiBADSinkhornMatcher(512).cuda().eval()
h1, h2 = O.texture_images(B, 480, 640, seed=5)
h1, h2 = h1.pin_memory(), h2.pin_memory()
u1, u2 = h1.to(torch.uint8).pin_memory(), h2.to(torch.uint8).pin_memory()
st = torch.cuda.current_stream()
wrapped = om.MatchExtractionWrapper(model, max_matches=100, match_threshold=0.0035).cuda().eval()
for name, a1, a2, mdl in (("f32", h1, h2, model), ("u8", u1, u2, model), ("u8-matches", u1, u2, wrapped)):
    for chunk, ns in ((8, 4), (16, 4), (16, 6), (32, 2), (32, 4), (64, 2), (64, 3), (64, 4)):
        hb = HostBatchMatcher(mdl, chunk=chunk, n_streams=ns, depth=max(2, ns), join=False)
        for _ in range(3): hb(a1, a2)
        hb.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(st)
        steps = 12
        for _ in range(steps): hb(a1, a2)
        for s in hb.streams: st.wait_stream(s)
        t1.record(st); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / steps
        print(f"{name} chunk {chunk:3d} streams {ns}: {ms:6.3f} ms/step  {B / ms * 1e3:8.0f} pairs/s")
