#!/usr/bin/env python3
"""A/B of the whole matcher step under each detector routing (0 default, 3 fused sweep, 4 split sweep, 2 tiled).
usage: python tools/ab_step.py [dense|sparse|angle] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import onnx_image_processing_b200 as om
from onnx_image_processing_b200 import _native as nat
from oracle import oracle as O   # synthetic inputs only

wl = sys.argv[1] if len(sys.argv) > 1 else "dense"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cls = {"dense": om.ShiTomasiBADSinkhornMatcher, "sparse": om.ShiTomasiSparseBADSinkhornMatcher,
       "angle": om.ShiTomasiAngleSparseBADSinkhornMatcher}[wl]
model = cls(512).cuda().eval()
sets = [tuple(t.cuda() for t in O.texture_images(B, 480, 640, seed=s)) for s in (3, 4, 5, 6)]   # 4 x 157 MB > L2
lib = nat.lib()
def run(steps):
    with torch.no_grad():
        for i in range(3): model(*sets[i % 4])
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for i in range(steps): model(*sets[i % 4])
        b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps
for rep in range(2):
    for mode in (0, 3, 4, 2):
        lib.om_debug_force_generic_stencil(mode)
        print(f"{wl} B={B} stencil routing {mode}: {run(20) * 1000:.1f} us/step")
lib.om_debug_force_generic_stencil(0)
for rep in range(3):
    for n in (1, 2, 4):
        lib.om_debug_match_streams(n)
        print(f"{wl} B={B} image chains on {n} stream(s): {run(30) * 1000:.1f} us/step")
lib.om_debug_match_streams(4)
