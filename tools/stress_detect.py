import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import oracle as O
from onnx_image_processing_b200 import _ops
cfgs = [dict(B=1, H=75, W=108, K=200, bs=3, r=1, margin=0, thr=500.0, seed=268546),
        dict(B=3, H=120, W=160, K=300, bs=3, r=3, margin=7, thr=0.0, seed=5),
        dict(B=2, H=97, W=333, K=513, bs=5, r=5, margin=0, thr=0.0, seed=6),
        dict(B=2, H=200, W=260, K=64, bs=5, r=3, margin=3, thr=0.0, seed=7),
        dict(B=8, H=480, W=640, K=512, bs=3, r=3, margin=7, thr=0.0, seed=8)]
for c in cfgs:
    img = O.texture_images(c["B"], c["H"], c["W"], seed=c["seed"])[0].cuda()
    k0, s0 = _ops.detect(img, c["K"], c["bs"], c["r"], c["thr"], c["margin"])
    bad = 0
    for it in range(400):
        # interleave a different shape so that workspaces / kernels alternate as in the test-suite
        if it % 3 == 0:
            _ops.detect(img[:, :, : c["H"] // 2, : c["W"] // 2].contiguous(), 7, 3, 3, 0.0, 0)
        k, s = _ops.detect(img, c["K"], c["bs"], c["r"], c["thr"], c["margin"])
        if not (torch.equal(k, k0) and torch.equal(s, s0)):
            bad += 1
            if bad <= 3:
                d = (k != k0).any(-1).nonzero().tolist()
                print("  iteration", it, "differs in", len(d), "slots, first", d[:4])
    print(c, "-> nondeterministic runs:", bad, "of 400")
