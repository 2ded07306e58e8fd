"""Loading of the committed golden vectors (tests/golden/*.npz, minted by tests/golden/make_golden.py)."""
import ast
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(kind=None):
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        n = os.path.splitext(os.path.basename(p))[0]
        if kind is None or str(np.load(p)["kind"]) == kind:
            out.append(n)
    return out


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {}
    for k in z.files:
        v = z[k]
        if v.dtype.kind in "US":
            g[k] = str(v)
        elif v.ndim == 0:
            g[k] = v.item()
        else:
            g[k] = torch.from_numpy(v)
    if "kwargs" in g:
        g["kwargs"] = ast.literal_eval(g["kwargs"])
    for k in ("image1", "image2"):
        if k in g:
            g[k] = g[k].float()
    return g
