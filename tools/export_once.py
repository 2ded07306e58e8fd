"""One step of the export-default configuration (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import onnx_image_processing_b200 as om
from oracle import oracle as O
base1, base2 = O.texture_images(64, 480, 640, seed=1)
m = om.ShiTomasiSparseBADSinkhornMatcher(1024, num_pairs=512, binarize=True, soft_binarize=False, epsilon=0.05, nms_radius=5).cuda().eval()
i1, i2 = base1.cuda(), base2.cuda()
with torch.no_grad():
    for _ in range(2):
        m(i1, i2)
torch.cuda.synchronize()
