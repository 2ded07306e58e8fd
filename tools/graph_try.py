#!/usr/bin/env python3
"""CUDA-graph capture of one matcher step (static inputs): replay latency vs eager calls."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import onnx_image_processing_b200 as om
from oracle import oracle as O
for wl, cls in (("sparse", om.ShiTomasiSparseBADSinkhornMatcher), ("dense", om.ShiTomasiBADSinkhornMatcher)):
    for B in (1, 8, 64):
        model = cls(512).cuda().eval()
        i1, i2 = (t.cuda() for t in O.texture_images(B, 480, 640, seed=3))
        with torch.no_grad():
            for _ in range(3): ref = model(i1, i2)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                model(i1, i2)
                with torch.cuda.graph(g, stream=s):
                    out = model(i1, i2)
            torch.cuda.current_stream().wait_stream(s)
            g.replay(); torch.cuda.synchronize()
            same = all(torch.equal(a, b) for a, b in zip(ref, out))
            def t(fn, n=50):
                for _ in range(5): fn()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); a.record()
                for _ in range(n): fn()
                b.record(); torch.cuda.synchronize()
                return a.elapsed_time(b) / n * 1000
            print(f"{wl} B={B}: eager {t(lambda: model(i1, i2)):.1f} us, graph replay {t(g.replay):.1f} us, identical {same}")
