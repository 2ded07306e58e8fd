"""Parity metrics shared by the GPU tests and tools/gpu_diag.py.

Tolerances are the north star's (BASELINE.json): keypoints identical except exact score ties;
descriptors <= 1e-5 abs (L2-normalised output; raw output relative to its magnitude);
match probabilities <= 1e-4 abs on the core block and the dustbin row/column, relative 1e-4 on
the dustbin/dustbin corner (a value of ~240-510); row-argmax agreement >= 99.9 %.
"""
from __future__ import annotations

import torch

DESC_TOL = 1e-5
# soft binarisation is sigmoid(-10 * centered): it multiplies the reference's own conv2d rounding
# noise on the box means (up to 2e-4 abs) by 2.5, so agreement with ANY exact
# box-mean implementation is limited to ~5e-5 on the normalised descriptor
DESC_TOL_SOFT = 1e-4
PROB_TOL = 1e-4
ARGMAX_MIN = 0.999
SCORE_REL_TOL = 4e-7      # keypoint scores: the reference's MKL sqrt is off by <= 1 ulp of the sqrt term


def scores_close(s_test: torch.Tensor, s_ref: torch.Tensor) -> bool:
    s_test, s_ref = s_test.cpu(), s_ref.cpu()
    return bool(((s_test - s_ref).abs() <= SCORE_REL_TOL * 64 * s_ref.abs().clamp_min(1.0)).all()) and \
        bool(((s_test > 0) == (s_ref > 0)).all())


def keypoint_mismatches(k_test: torch.Tensor, k_ref: torch.Tensor, s_ref: torch.Tensor | None = None) -> int:
    """Number of keypoint slots that differ and are not excused by an exact score tie."""
    k_test, k_ref = k_test.cpu(), k_ref.cpu()
    diff = (k_test != k_ref).any(dim=-1)
    if s_ref is None or not diff.any():
        return int(diff.sum())
    s_ref = s_ref.cpu()
    bad = 0
    for b in range(k_ref.shape[0]):
        idx = diff[b].nonzero().flatten()
        for i in idx.tolist():
            tied = (s_ref[b] == s_ref[b, i]).sum() > 1 and s_ref[b, i] > 0
            if not tied:
                bad += 1
            else:
                # the tied group must hold the same SET of coordinates
                grp = (s_ref[b] == s_ref[b, i])
                a = {tuple(v) for v in k_test[b][grp].tolist()}
                r = {tuple(v) for v in k_ref[b][grp].tolist()}
                bad += int(a != r)
    return bad


def desc_metrics(d_test: torch.Tensor, d_ref: torch.Tensor, tol: float = DESC_TOL) -> dict:
    d_test, d_ref = d_test.cpu(), d_ref.cpu()
    err = (d_test - d_ref).abs()
    scale = d_ref.abs().amax().clamp_min(1.0)
    row_ok = (err.amax(dim=-1) <= tol * scale)
    return dict(max_abs=float(err.max()), max_rel_to_scale=float(err.max() / scale),
                rows_within=float(row_ok.float().mean()), elems_over=float((err > tol * scale).float().mean()))


def prob_metrics(p_test: torch.Tensor, p_ref: torch.Tensor) -> dict:
    p_test, p_ref = p_test.cpu(), p_ref.cpu()
    N, M = p_ref.shape[1] - 1, p_ref.shape[2] - 1
    err = (p_test - p_ref).abs()
    core = float(err[:, :N, :].max()) if N > 0 else 0.0
    dust_row = float(err[:, N, :M].max()) if M > 0 else 0.0
    corner_rel = float((err[:, N, M] / p_ref[:, N, M].abs().clamp_min(1e-30)).max())
    am_t = p_test[:, :N, :].argmax(dim=-1)
    am_r = p_ref[:, :N, :].argmax(dim=-1)
    return dict(core=core, dust_row=dust_row, corner_rel=corner_rel, argmax=float((am_t == am_r).float().mean()),
                finite=bool(torch.isfinite(p_test).all()))


def probs_ok(m: dict) -> bool:
    return (m["finite"] and m["core"] <= PROB_TOL and m["dust_row"] <= PROB_TOL and m["corner_rel"] <= PROB_TOL
            and m["argmax"] >= ARGMAX_MIN)


def p_summary_metrics(p: torch.Tensor, g: dict) -> dict:
    """(1, N+1, M+1) matrix against the row-sample form stored for the large goldens (tests/golden/make_golden_full.py):
    sampled full rows, the complete dustbin row / column, every row's maximum and argmax, every column's argmax."""
    p = p.cpu()
    N = p.shape[1] - 1
    rows = g["P_rows"].long()
    core = p[0, :N, :N]
    corner_ref = float(g["P_dust_row"][N])
    return dict(core=max(float((p[0, rows] - g["P_sample"]).abs()[:, :N].max()),
                         float((core.max(dim=-1).values - g["P_row_max"]).abs().max()),
                         float((p[0, :N, N] - g["P_dust_col"][:N]).abs().max())),
                dust_row=float((p[0, N, :N] - g["P_dust_row"][:N]).abs().max()),
                corner_rel=abs(float(p[0, N, N]) - corner_ref) / max(abs(corner_ref), 1e-30),
                argmax=float((core.argmax(dim=-1) == g["P_row_argmax"]).float().mean()),
                col_argmax=float((core.argmax(dim=-2) == g["P_col_argmax"]).float().mean()),
                sum_rel=abs(float(p.double().sum()) - g["P_sum64"]) / abs(g["P_sum64"]),
                finite=bool(torch.isfinite(p).all()))


def p_summary_ok(p: torch.Tensor, g: dict, tol: float = PROB_TOL, argmax_min: float = 1.0) -> None:
    m = p_summary_metrics(p, g)
    assert m["finite"] and m["core"] <= tol and m["dust_row"] <= tol and m["corner_rel"] <= max(tol, PROB_TOL), m
    assert m["argmax"] >= argmax_min and m["col_argmax"] >= argmax_min and m["sum_rel"] <= 1e-5, m


def full_descriptor_bits(g: dict, which: int) -> torch.Tensor:
    """hard-binarised descriptors stored as packed bits + the row norm factor -> float descriptors"""
    import numpy as np
    bits = torch.from_numpy(np.unpackbits(g[f"desc{which}_bits"].numpy(), axis=-1)).float()
    return bits * g[f"desc{which}_norm"].unsqueeze(-1)
