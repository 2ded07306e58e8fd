#!/usr/bin/env python3
"""Score / NMS / top-k kernel times against the batch size (is the split detector bound by DRAM or by latency?)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_image_processing_b200 import _native as nat
from oracle import oracle as O
lib = nat.lib()
dev = "cuda:0"
H, W, K = 480, 640, 512
base, _ = O.texture_images(64, H, W, seed=1)
st = torch.cuda.current_stream()
sp = ctypes.c_void_p(st.cuda_stream)
ptr = lambda t: ctypes.c_void_p(t.data_ptr())
for B in (8, 16, 32, 64, 128, 256):
    img = torch.cat([base] * ((B + 63) // 64))[:B].contiguous().to(dev)
    kp = torch.empty((B, K, 2), device=dev); ks = torch.empty((B, K), device=dev)
    ws = torch.empty(lib.om_topk_workspace_bytes(B, H, W, K), dtype=torch.uint8, device=dev)
    out = []
    for stage in (2, 3, 1):
        f = lambda: nat.check(lib.om_debug_detect_stage(ptr(img), B, H, W, 3, 3, 7, 0.0, K, ptr(kp), ptr(ks), ptr(ws), ws.numel(), sp, stage), "stage")
        for _ in range(3): f()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(20): f()
        b.record(); torch.cuda.synchronize()
        out.append(a.elapsed_time(b) / 20 * 1e3)
    print(f"B={B:4d}: score {out[0]:7.1f} us ({out[0]/B:5.2f}/img)  nms {out[1]:7.1f} us ({out[1]/B:5.2f}/img)  topk {out[2]:7.1f} us ({out[2]/B:5.2f}/img)")
