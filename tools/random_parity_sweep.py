#!/usr/bin/env python3
"""Many seeded random detector cases against the oracle (hunting rare mismatches).  usage: random_parity_sweep.py [n] [seed]"""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import oracle as O
from onnx_image_processing_b200 import _ops
from tests import parity as PR
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
for ci in range(n):
    H, W = rnd.randint(24, 200), rnd.randint(24, 260)
    c = dict(B=rnd.randint(1, 3), H=H, W=W, K=rnd.choice([1, 7, 64, 200, 513]), bs=rnd.choice([3, 5]), r=rnd.choice([1, 2, 3, 5]),
             margin=rnd.choice([0, 0, 3, 7, 11]), thr=rnd.choice([0.0, 0.0, 500.0]), family=rnd.choice(["texture", "texture", "noise"]),
             seed=rnd.randint(0, 10 ** 6))
    if c["K"] > H * W:
        continue
    img = (O.texture_images(c["B"], H, W, seed=c["seed"])[0] if c["family"] == "texture" else O.noise_images(c["B"], H, W, seed=c["seed"]))
    rk, rs = O.detect(img, c["K"], c["bs"], c["r"], c["thr"], c["margin"])
    for rep in range(3):
        gk, gs = _ops.detect(img.cuda(), c["K"], c["bs"], c["r"], c["thr"], c["margin"])
        m = PR.keypoint_mismatches(gk, rk, rs)
        if m:
            bad += 1
            d = (gk.cpu() != rk).any(-1).nonzero().tolist()
            print("MISMATCH case", ci, "rep", rep, c, m, [(b, i, rk[b, i].tolist(), float(rs[b, i]), gk[b, i].tolist(), float(gs[b, i])) for b, i in d[:4]])
print("cases", n, "mismatching runs", bad)
