#!/usr/bin/env python3
"""Phase timing of the tcgen05 Sinkhorn kernel from in-kernel clock64 stamps (debug aid)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import onnx_image_processing_b200 as om
from onnx_image_processing_b200 import _native as nat

B = int(sys.argv[1]) if len(sys.argv) > 1 else 15
g = torch.Generator().manual_seed(0)
d1 = torch.nn.functional.normalize(torch.randn(B, 512, 256, generator=g), dim=-1).cuda()
d2 = torch.nn.functional.normalize(torch.randn(B, 512, 256, generator=g), dim=-1).cuda()
m = om.SinkhornMatcher(20, 1.0).cuda()
for _ in range(3):
    m(d1, d2)
buf = torch.zeros(B * 8 * 16, dtype=torch.int64, device="cuda")
nat.lib().om_debug_sinkhorn_trace(ctypes.c_void_p(buf.data_ptr()))
m(d1, d2)
torch.cuda.synchronize()
nat.lib().om_debug_sinkhorn_trace(ctypes.c_void_p(0))
t = buf[:B * 64].view(B * 8, 8).cpu().double()
x = buf[B * 64:B * 96].view(B * 8, 4).cpu().double() / 20.0
y = buf[B * 96:].view(B * 8, 4).cpu().double() / 8.0
names = ["staging loop (16 chunks)", "wait last MMAs", "epilogue TMEM->smem", "init u,v + cluster.sync", "20 iterations", "write P"]
for i, n in enumerate(names):
    d = t[:, i + 1] - t[:, i]
    print(f"{n:28s} mean {d.mean():9.0f} cyc  min {d.min():9.0f}  max {d.max():9.0f}")
print(f"{'  of which waiting partials':28s} mean {t[:, 7].mean():9.0f} cyc  min {t[:,7].min():9.0f} max {t[:,7].max():9.0f}")
print(f"{'total':28s} mean {(t[:, 6] - t[:, 0]).mean():9.0f} cyc")
for i, n in enumerate(["sweep (thread 0 view)", "barrier+reduce+send", "wait partials", "owner + wait b"]):
    print(f"  per iteration {n:24s} mean {x[:, i].mean():8.0f} cyc  min {x[:, i].min():8.0f}  max {x[:, i].max():8.0f}")
for i, n in enumerate(["wait stage free", "split + store + arrive + fetch", "wait stage full", "MMA issue (thread 0)"]):
    print(f"  per chunk (thread 0) {n:24s} mean {y[:, i].mean():8.0f} cyc  min {y[:, i].min():8.0f}  max {y[:, i].max():8.0f}")
