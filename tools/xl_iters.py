"""Per-iteration cost of the streaming Sinkhorn path: the stage timed at several iteration counts."""
import sys, torch
sys.path.insert(0, ".")
from onnx_image_processing_b200 import _native, _ops
lib = _native.lib()
B, K = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 2048
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
lib.om_debug_xl_reverse(mode)
g = torch.Generator().manual_seed(1)
d1 = torch.nn.functional.normalize(torch.randn(B, K, 256, generator=g), dim=-1).cuda()
d2 = torch.nn.functional.normalize(d1 + 0.2 * torch.randn(B, K, 256, device="cuda"), dim=-1)
res = {}
for its in (1, 20, 40, 100):
    for _ in range(2):
        _ops.sinkhorn(d1, d2, its, 1.0, 1.0, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _ops.sinkhorn(d1, d2, its, 1.0, 1.0, False)
    e1.record()
    torch.cuda.synchronize()
    res[its] = e0.elapsed_time(e1) / 5 * 1000
    print(f"B={B} K={K} mode={mode} iterations={its}: {res[its]:.1f} us", flush=True)
print(f"per iteration: {(res[100] - res[20]) / 80:.2f} us; fixed part: {res[20] - 20 * (res[100] - res[20]) / 80:.1f} us")
