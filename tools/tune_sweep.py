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

# kernel variants of the split form (score variant, NMS variant): 0 = generic sweep kernels, 1 = lean kernels at 5 CTAs/SM, 2 = at 6
def t_run(n=20):
    for _ in range(3): run()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): run()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1000
lib.om_debug_force_generic_stencil(4)
for sv, nv in ((0, 0), (0, 1), (1, 1), (2, 1), (1, 2), (2, 2)):
    lib.om_debug_score_variant(sv); lib.om_debug_nms_variant(nv)
    for strip_a, strip_b in ((24, 32), (32, 32), (40, 32), (48, 32), (60, 32), (32, 64), (60, 64)):
        lib.om_debug_sweep_tuning(strip_a, 100 + strip_b)
        print(f"score variant {sv}, NMS variant {nv}, strips {strip_a}/{strip_b}: {t_run():.1f} us")
lib.om_debug_score_variant(1); lib.om_debug_nms_variant(1)
lib.om_debug_sweep_tuning(0, 0)
lib.om_debug_force_generic_stencil(0)
# default routing
for _ in range(3): run()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(20): run()
b.record(); torch.cuda.synchronize()
print(f"default routing: {a.elapsed_time(b) / 20 * 1000:.1f} us")
