#!/usr/bin/env python3
"""Measured parity figures of the CUDA path on the golden cases (GPU box): what the thresholds in
tests/test_gpu_parity.py are set from.  Writes one JSON object (default gpurun_out/parity_measured.json)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import onnx_image_processing_b200 as om  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests import golden_util as G  # noqa: E402
from tests import parity as PR  # noqa: E402

DEV = "cuda:0"


def cuda(*ts):
    return [t.to(DEV) for t in ts]


def oriented_rows(d, ref, theta_gpu=None, theta_ref=None):
    """rows beyond 1e-5, their cosine to the reference and the orientation difference at those keypoints"""
    d, ref = d.cpu(), ref.cpu()
    err = (d - ref).abs().amax(dim=-1)
    bad = err > PR.DESC_TOL
    out = dict(rows=int(bad.numel()), rows_over=int(bad.sum()), frac_over=float(bad.float().mean()),
               max_abs_good_rows=float(err[~bad].max()) if (~bad).any() else 0.0)
    if bad.any():
        cos = torch.nn.functional.cosine_similarity(d[bad], ref[bad], dim=-1)
        out["min_cos_over"] = float(cos.min())
        if theta_gpu is not None:
            dt = (theta_gpu.cpu() - theta_ref.cpu()).abs()
            dt = torch.minimum(dt, (2 * torch.pi - dt).abs())
            out["max_dtheta_over"] = float(dt[bad].max())
            out["max_dtheta_all"] = float(dt.max())
    return out


def theta_at(model, img, kpts):
    a = model.detector.angle_estimator(img.to(DEV))[:, 0]
    yi = kpts[..., 0].clamp(min=0).long().to(DEV)
    xi = kpts[..., 1].clamp(min=0).long().to(DEV)
    return a[torch.arange(a.shape[0], device=DEV)[:, None], yi, xi]


def main():
    out = {}
    with torch.no_grad():
        g = G.load("dense_full_default")
        k1, k2, p, d1, d2 = om.ShiTomasiBADSinkhornMatcher(512).to(DEV).eval().match(*cuda(g["image1"], g["image2"]))
        out["dense_full_default"] = dict(kp1=PR.keypoint_mismatches(k1, g["kpts1"]), kp2=PR.keypoint_mismatches(k2, g["kpts2"]),
                                         desc1=PR.desc_metrics(d1, g["desc1"]), desc2=PR.desc_metrics(d2, g["desc2"]),
                                         P=PR.prob_metrics(p, g["P"]))

        g = G.load("sparse_full_export")
        m = om.ShiTomasiSparseBADSinkhornMatcher(1024, **g["kwargs"]).to(DEV).eval()
        k1, k2, p, d1, d2 = m.match(*cuda(g["image1"], g["image2"]))
        r1, r2 = PR.full_descriptor_bits(g, 1), PR.full_descriptor_bits(g, 2)
        flips = [int(((a.cpu() > 0) != (b > 0)).sum()) for a, b in ((d1, r1), (d2, r2))]
        rows = [int(((a.cpu() > 0) != (b > 0)).any(dim=-1).sum()) for a, b in ((d1, r1), (d2, r2))]
        out["sparse_full_export"] = dict(kp1=PR.keypoint_mismatches(k1, g["kpts1"]), kp2=PR.keypoint_mismatches(k2, g["kpts2"]),
                                         flipped_bits=flips, rows_with_flips=rows, P=PR.p_summary_metrics(p, g))

        g = G.load("sparse_1080p_k2048")
        k1, k2, p, d1, d2 = om.ShiTomasiSparseBADSinkhornMatcher(2048).to(DEV).eval().match(*cuda(g["image1"], g["image2"]))
        rows = g["desc_rows"].long()
        out["sparse_1080p_k2048"] = dict(kp1=PR.keypoint_mismatches(k1, g["kpts1"]), kp2=PR.keypoint_mismatches(k2, g["kpts2"]),
                                         desc1=PR.desc_metrics(d1[0, rows], g["desc1_sample"]),
                                         desc2=PR.desc_metrics(d2[0, rows], g["desc2_sample"]),
                                         rowsum1=float((d1[0].double().sum(-1).float().cpu() - g["desc1_rowsum"]).abs().max()),
                                         rowsum2=float((d2[0].double().sum(-1).float().cpu() - g["desc2_rowsum"]).abs().max()),
                                         P=PR.p_summary_metrics(p, g))

        for name in G.names("angle"):
            g = G.load(name)
            m = om.ShiTomasiAngleSparseBADSinkhornMatcher(g["K"], **g["kwargs"]).to(DEV).eval()
            k1, k2, p, d1, d2 = m.match(*cuda(g["image1"], g["image2"]))
            th = theta_at(m, g["image1"], g["kpts1"])
            out[name] = dict(kp1=PR.keypoint_mismatches(k1, g["kpts1"]), kp2=PR.keypoint_mismatches(k2, g["kpts2"]),
                             desc1=oriented_rows(d1, g["desc1"], th, g["theta1_at_kpts"]), desc2=oriented_rows(d2, g["desc2"]),
                             P=PR.prob_metrics(p, g["P"]))

        for name in G.names("sparse"):
            g = G.load(name)
            kw = g["kwargs"]
            if not (kw.get("binarize") and not kw.get("soft_binarize", True)):
                continue
            m = om.ShiTomasiSparseBADSinkhornMatcher(g["K"], **kw).to(DEV).eval()
            k1, k2, p, d1, d2 = m.match(*cuda(g["image1"], g["image2"]))
            fl = ((d1.cpu() > 0) != (g["desc1"] > 0))
            raw = O.sparse_bad(g["image1"], g["kpts1"], None, num_pairs=kw.get("num_pairs", 256), normalize_descriptors=False)
            out[name] = dict(flipped_bits=int(fl.sum()), rows_with_flips=int(fl.any(-1).sum()), rows=int(fl.shape[0] * fl.shape[1]),
                             max_abs_centered_at_flips=float(raw[fl].abs().max()) if fl.any() else 0.0,
                             P=PR.prob_metrics(p, g["P"]))
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_measured.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
