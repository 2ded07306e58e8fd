"""One call of the Sinkhorn stage at K = 2048 (for an ncu launch list of the streaming path)."""
import sys, torch
sys.path.insert(0, ".")
from onnx_image_processing_b200 import _native, _ops
B, K = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 2048
g = torch.Generator().manual_seed(1)
d1 = torch.nn.functional.normalize(torch.randn(B, K, 256, generator=g), dim=-1).cuda()
d2 = torch.nn.functional.normalize(d1 + 0.2 * torch.randn(B, K, 256, device="cuda"), dim=-1)
for _ in range(2):
    p = _ops.sinkhorn(d1, d2, 20, 1.0, 1.0, False)
torch.cuda.synchronize()
