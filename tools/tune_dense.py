#!/usr/bin/env python3
"""Time the dense descriptor kernel (dense_at_kpts_kernel) alone: cp.async windows vs TMA boxes.
usage: python tools/tune_dense.py [batch]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import onnx_image_processing_b200 as om
from onnx_image_processing_b200 import _native as nat, _ops
from oracle import oracle as O   # synthetic inputs only
lib = nat.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H, W, K, P = 480, 640, 512, 256
img = O.texture_images(B, H, W, seed=3)[0].cuda()
model = om.ShiTomasiBADSinkhornMatcher(K).cuda().eval()
d = model.detector.descriptor
kp, _ = _ops.detect(img, K, 3, 3, 0.0, 0)
desc = torch.empty(B, K, P, device="cuda")
ws = torch.empty(lib.om_dense_bad_workspace_bytes(B, H, W), dtype=torch.uint8, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: ctypes.c_void_p(t.data_ptr())
def stage(s):
    nat.check(lib.om_debug_dense_stage(p(img), B, H, W, p(kp), K, p(d._pair_table), P, _ops.desc_mode(d.binarize, d.soft_binarize),
                                       float(d.temperature), 1, p(desc), p(ws), ws.numel(), st, s), "stage")
stage(0)
outs = []
for tma in (1, 0, 3, 2, 5, 4, 7, 6):
    lib.om_debug_dense_window(tma)
    for _ in range(3): stage(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(20): stage(1)
    b.record(); torch.cuda.synchronize()
    outs.append(desc.clone())
    print(f"{'TMA box' if tma & 1 else 'cp.async'} windows{', no fetch' if tma & 2 else ''}{', no arithmetic' if tma & 4 else ''}: {a.elapsed_time(b) / 20 * 1000:.1f} us")
lib.om_debug_dense_window(1)
print("identical:", torch.equal(outs[0], outs[1]))
