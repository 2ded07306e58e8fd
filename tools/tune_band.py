#!/usr/bin/env python3
"""Integral-image build (dense path, 64 images of 480x640) timed with CUDA events for several band heights."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import onnx_image_processing_b200 as om
from onnx_image_processing_b200 import _native as nat
from oracle import oracle as O

B, H, W, K, P = 64, 480, 640, 512, 256
dev = "cuda:0"
lib = nat.lib()
img = O.texture_images(B, H, W, seed=1000)[0].to(dev)
imgs = [img.clone() for _ in range(3)]                      # rotate inputs: 3 x 79 MB > L2
kp = torch.zeros((B, K, 2), device=dev)
desc = torch.empty((B, K, P), device=dev)
tb = om.BADDescriptor()._pair_table.to(dev)
ws = torch.empty(lib.om_dense_bad_workspace_bytes(B, H, W), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream()
sp = ctypes.c_void_p(st.cuda_stream)
p = lambda t: ctypes.c_void_p(t.data_ptr())
def run(i):
    nat.check(lib.om_debug_dense_stage(p(imgs[i % 3]), B, H, W, p(kp), K, p(tb), P, 0, 10.0, 1, p(desc), p(ws), ws.numel(), sp, 0), "stage")
for rows in (8, 12, 16, 24, 32, 48, 64, 128):
    lib.om_debug_band_rows(rows)
    for i in range(5): run(i)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for i in range(30): run(i)
    b.record(); torch.cuda.synchronize()
    print(f"band rows {rows:4d}: {a.elapsed_time(b) / 30 * 1e3:7.1f} us per 64 images (memset + colsum + band + 2 gated launches)")
