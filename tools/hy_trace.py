#!/usr/bin/env python3
"""Phase timing of the hybrid-resident Sinkhorn kernel (sinkhorn_hy.cu) from its in-kernel clock64 stamps (debug aid).
    python tools/hy_trace.py [B] [K] [D] [eps]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import onnx_image_processing_b200 as om
from onnx_image_processing_b200 import _native as nat

B = int(sys.argv[1]) if len(sys.argv) > 1 else 37
K = int(sys.argv[2]) if len(sys.argv) > 2 else 512
D = int(sys.argv[3]) if len(sys.argv) > 3 else 256
eps = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
CL = 4 if K <= 512 else 16
g = torch.Generator().manual_seed(0)
d1 = torch.nn.functional.normalize(torch.randn(B, K, D, generator=g), dim=-1).cuda()
d2 = torch.nn.functional.normalize(torch.randn(B, K, D, generator=g), dim=-1).cuda()
m = om.SinkhornMatcher(20, eps).cuda()
for _ in range(3):
    m(d1, d2)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    m(d1, d2)
e1.record()
torch.cuda.synchronize()
print(f"B={B} K={K} D={D} eps={eps}: {e0.elapsed_time(e1) * 100:.1f} us per call (pack kernels included)")
buf = torch.zeros(B * CL * 16, dtype=torch.int64, device="cuda")
nat.lib().om_debug_sinkhorn_trace(ctypes.c_void_p(buf.data_ptr()))
m(d1, d2)
torch.cuda.synchronize()
nat.lib().om_debug_sinkhorn_trace(ctypes.c_void_p(0))
t = buf[:B * CL * 8].view(B * CL, 8).cpu().double()
x = buf[B * CL * 8:].view(B * CL, 8).cpu().double() / 20.0
names = ["GEMM issue (thread 0)", "wait last MMAs", "epilogue TMEM->K", "init b + cluster.sync", "iterations", "P / fused epilogue"]
for i, n in enumerate(names):
    d = t[:, i + 1] - t[:, i]
    print(f"{n:28s} mean {d.mean():9.0f} cyc  min {d.min():9.0f}  max {d.max():9.0f}")
tot = t[:, 6] - t[:, 0]
print(f"{'total':28s} mean {tot.mean():9.0f} cyc  min {tot.min():9.0f}  max {tot.max():9.0f}")
print(f"kernel span (clock64 is per SM; indicative): {(t[:, 6].max() - t[:, 0].min()):.0f} cyc")
if x.abs().sum() > 0:      # only builds that carry per-iteration stamps fill this half of the buffer
    for i, n in enumerate(["phase 0", "phase 1", "phase 2", "phase 3", "phase 4", "phase 5"]):
        print(f"  per iteration {n:26s} mean {x[:, i].mean():8.0f} cyc  min {x[:, i].min():8.0f}  max {x[:, i].max():8.0f}")
