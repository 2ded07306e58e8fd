#!/usr/bin/env python3
"""Time the detector stencil kernel for a few (strip rows, CTAs/SM) settings."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_image_processing_b200 import _native as nat
from oracle import oracle as O
lib = nat.lib()
B, H, W, K = 64, 480, 640, 512
img, _ = O.texture_images(B, H, W, seed=3)
img = img.cuda()
nat.use_device(0)
ws = torch.empty(lib.om_topk_workspace_bytes(B, H, W, K), dtype=torch.uint8, device="cuda")
kp = torch.empty(B, K, 2, device="cuda"); ks = torch.empty(B, K, device="cuda")
st = torch.cuda.current_stream()
p = lambda t: ctypes.c_void_p(t.data_ptr())
BS, R = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3, 3)
def run():
    nat.check(lib.om_debug_detect_stage(p(img), B, H, W, BS, R, 7, 0.0, K, p(kp), p(ks), p(ws), ws.numel(), ctypes.c_void_p(st.cuda_stream), 0), "stage")
print("block", BS, "radius", R)
for strip, minb in [(40, 3), (40, 4), (60, 3), (30, 4), (48, 4), (24, 99), (16, 99), (20, 99), (16, 124), (24, 116), (24, 132), (32, 124), (12, 124), (24, 140)]:   # 99 = split score/NMS kernels
    lib.om_debug_sweep_tuning(strip, minb)
    lib.om_debug_force_generic_stencil(4 if minb >= 99 else 3)
    for _ in range(3): run()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(20): run()
    b.record(); torch.cuda.synchronize()
    print(f"strip {strip:4d} minb {minb}: {a.elapsed_time(b) / 20 * 1000:.1f} us")
lib.om_debug_sweep_tuning(0, 0)
for mode, name in ((2, "tiled shared-memory kernel"), (1, "generic kernel")):
    lib.om_debug_force_generic_stencil(mode)
    for _ in range(3): run()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(10): run()
    b.record(); torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b) / 10 * 1000:.1f} us")
lib.om_debug_force_generic_stencil(0)

# NMS kernel variants of the split form
for v, name in ((0, "any-radius nms_sweep_kernel"), (2, "nms3 at 6 CTAs/SM"), (1, "nms3 at 5 CTAs/SM (default)")):
    lib.om_debug_nms_variant(v)
    for strip_b in (24, 32, 48, 64):
        lib.om_debug_sweep_tuning(24, 100 + strip_b)
        for _ in range(3): run()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(20): run()
        b.record(); torch.cuda.synchronize()
        print(f"{name}, NMS strip {strip_b}: {a.elapsed_time(b) / 20 * 1000:.1f} us")
lib.om_debug_sweep_tuning(0, 0)
# default routing
for _ in range(3): run()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(20): run()
b.record(); torch.cuda.synchronize()
print(f"default routing: {a.elapsed_time(b) / 20 * 1000:.1f} us")
