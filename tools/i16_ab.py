"""A/B of the sparse descriptor path's integral: modulo 2^16 (default) against uint32 (om_debug_band_rows(-32)); device-resident
steps of the sparse default, export-default and 1080p configurations, CUDA events."""
import sys, torch
sys.path.insert(0, ".")
import onnx_image_processing_b200 as om
from onnx_image_processing_b200 import _native
from oracle import oracle as O
lib = _native.lib()
i1, i2 = (t.cuda() for t in O.texture_images(64, 480, 640, seed=1))
b1, b2 = (t.cuda() for t in O.texture_images(8, 1080, 1920, seed=2))
cases = {"sparse": (om.ShiTomasiSparseBADSinkhornMatcher(512), i1, i2),
         "angle": (om.ShiTomasiAngleSparseBADSinkhornMatcher(512), i1, i2),
         "export": (om.ShiTomasiSparseBADSinkhornMatcher(1024, num_pairs=512, binarize=True, soft_binarize=False, epsilon=0.05, nms_radius=5), i1, i2),
         "1080p": (om.ShiTomasiSparseBADSinkhornMatcher(2048), b1, b2)}
for name, (m, x1, x2) in cases.items():
    m = m.cuda().eval()
    for mode in (-16, -32, -16, -32):
        lib.om_debug_band_rows(mode)
        with torch.no_grad():
            for _ in range(5):
                out = m(x1, x2)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                out = m(x1, x2)
            b.record()
            torch.cuda.synchronize()
        print(f"{name} integral bits={-mode}: {a.elapsed_time(b) / 20:.4f} ms", flush=True)
lib.om_debug_band_rows(-16)
