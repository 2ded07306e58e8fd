#!/usr/bin/env python3
"""Live check of the oracle against the reference itself (authoring container only: needs /root/reference).

    PYTHONDONTWRITEBYTECODE=1 python oracle/validate_against_reference.py [--seeds N]

For several seeds and input families it runs the reference's unified modules and the oracle's restatements on the
same inputs and requires bit-identical keypoints and probabilities for the sparse and oriented matchers; for the
dense matcher the dense map and keypoints are bit-identical and P agrees to 2e-6 (the oracle gathers the (B,256,H,W)
map pair-chunk by pair-chunk to bound memory, which changes grid_sample's vectorisation remainder handling).  tests/golden/make_golden.py commits
a subset of exactly these outputs as fixtures so that the same check runs where the reference is absent.
TEST INFRASTRUCTURE ONLY.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("OM_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import torch  # noqa: E402

from oracle import oracle as O  # noqa: E402


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=3)
    ap.add_argument("--size", type=int, nargs=2, default=[120, 160])
    args = ap.parse_args()
    if not os.path.isdir(REF):
        print(f"reference not found at {REF}: nothing to validate against")
        return 0
    from pytorch_model.feature_detection.shi_tomasi_sparse_bad_sinkhorn import ShiTomasiSparseBADSinkhornMatcher
    from pytorch_model.feature_detection.shi_tomasi_angle_sparse_bad_sinkhorn import ShiTomasiAngleSparseBADSinkhornMatcher
    from pytorch_model.feature_detection.shi_tomasi_bad_sinkhorn import ShiTomasiBADSinkhornMatcher

    H, W = args.size
    K = 96
    bad = 0
    with torch.no_grad():
        for seed in range(args.seeds):
            for fam, gen in (("texture", O.texture_images), ("noise", O.noise_images)):
                i1, i2 = gen(2, H, W, seed=seed)[:2] if fam == "texture" else (gen(2, H, W, seed=seed), gen(2, H, W, seed=seed + 100))
                cases = [
                    ("sparse", ShiTomasiSparseBADSinkhornMatcher(K).eval(), lambda a, b: O.sparse_matcher(a, b, K)),
                    ("angle", ShiTomasiAngleSparseBADSinkhornMatcher(K).eval(), lambda a, b: O.angle_matcher(a, b, K)),
                    ("dense", ShiTomasiBADSinkhornMatcher(K).eval(), lambda a, b: O.dense_matcher(a, b, K)),
                ]
                for name, ref_mod, ora in cases:
                    rk1, rk2, rp = ref_mod(i1, i2)
                    ok1, ok2, op = ora(i1, i2)[:3]
                    err = float((rp - op)[:, :K, :].abs().max())
                    kp_same = torch.equal(rk1, ok1) and torch.equal(rk2, ok2)
                    same = kp_same and (torch.equal(rp, op) if name != "dense" else err <= 2e-6)
                    what = "bit-identical" if torch.equal(rp, op) and kp_same else ("keypoints identical, P within 2e-6" if same else "DIFFERENT")
                    print(f"seed {seed} {fam:8s} {name:7s}: {what} (max |dP| {err:.3g})")
                    bad += 0 if same else 1
    print("oracle == reference on every case" if bad == 0 else f"{bad} case(s) differ")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
