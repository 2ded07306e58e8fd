import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import onnx_image_processing_b200 as om
from onnx_image_processing_b200 import _native as nat
from oracle import oracle as O
lib = nat.lib(); dev = "cuda:0"
img, _ = O.texture_images(2, 150, 210, seed=7)
d = img.to(dev)
B, H, W, K = 2, 150, 210, 300
kp = torch.empty((B, K, 2), device=dev); ks = torch.empty((B, K), device=dev)
sm = torch.zeros((B, H, W), device=dev)
ws = torch.empty(lib.om_topk_workspace_bytes(B, H, W, K), dtype=torch.uint8, device=dev)
p = lambda t: ctypes.c_void_p(t.data_ptr())
nat.check(lib.om_detect_f32(p(d), B, H, W, 3, 3, 4, 0.0, K, p(sm), p(kp), p(ks), p(ws), ws.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "detect")
torch.cuda.synchronize()
ref = O.shi_tomasi_score(img, 3, ieee_sqrt=True).squeeze(1)
got = sm.cpu()
bad = (got != ref)
print("mismatches", int(bad.sum()), "of", bad.numel())
idx = bad.nonzero()[:12]
for b, y, x in idx.tolist():
    print(b, y, x, float(got[b, y, x]), float(ref[b, y, x]), float(got[b, y, x] - ref[b, y, x]))
print("cols with mismatch", sorted(set(idx[:, 2].tolist()))[:20], "rows", sorted(set(bad.nonzero()[:, 1].tolist()))[:20])
