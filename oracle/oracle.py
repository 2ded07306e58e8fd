"""CPU oracle for the feature-matching hot path.  TEST INFRASTRUCTURE ONLY -- never shipped,
never on the product path (see oracle/__init__.py).

Every function restates one stage of the reference with the same ATen CPU operators in the
same order, so that rounding behaviour (oneDNN conv, double-accumulated cumsum, partial_sort
top-k, grid_sampler, max-shifted logsumexp) is the reference's own.  Citations are
file:line under the reference checkout (fateshelled/onnx_image_processing).

PARITY PIN: the reference ships no golden vectors for this path (its only tests have no
asserts).  The oracle is pinned instead against outputs of the reference itself, run in the
authoring container: tests/golden/make_golden.py imports the reference, runs the unified
modules on seeded inputs and commits the results under tests/golden/; tests/test_oracle_golden.py
checks this file against them bit-for-bit.  oracle/validate_against_reference.py repeats the
comparison live on many more seeds when /root/reference is present.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch
import torch.nn.functional as F

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                     "onnx_image_processing_b200", "data", "bad_tables.npz")


# --------------------------------------------------------------------------------------
# constant tables (descriptor/bad_params.py:4-1568, data only)
# --------------------------------------------------------------------------------------
def bad_tables(num_pairs: int):
    """(ox1, ox2, oy1, oy2) float32 offsets relative to the keypoint, int64 radii, float32 thresholds.

    Offsets are table-16 as in descriptor/bad.py:33-37 / :405-410.
    """
    if num_pairs not in (256, 512):
        raise ValueError(f"num_pairs must be 256 or 512 to use learned BAD patterns, got {num_pairs}")
    z = np.load(_DATA)
    box = torch.from_numpy(z[f"boxes{num_pairs}"].astype(np.float32))
    thr = torch.from_numpy(z[f"thresholds{num_pairs}"].astype(np.float32))
    return box[:, 0].clone(), box[:, 1].clone(), box[:, 2].clone(), box[:, 3].clone(), box[:, 4].to(torch.int64), thr


def box_kernel_bank(max_radius: int = 7) -> torch.Tensor:
    """(R+1,1,2R+1,2R+1) normalised square masks, descriptor/bad.py:426-434."""
    c = torch.arange(-max_radius, max_radius + 1, dtype=torch.float32)
    gy, gx = torch.meshgrid(c, c, indexing="ij")
    rv = torch.arange(max_radius + 1, dtype=torch.float32).view(-1, 1, 1)
    masks = ((gy.abs() <= rv) & (gx.abs() <= rv)).to(torch.float32)
    denom = ((2.0 * rv + 1.0) ** 2).clamp_min(1.0)
    return (masks / denom).unsqueeze(1)


# --------------------------------------------------------------------------------------
# a1: Shi-Tomasi score (detector/shi_tomasi.py:66-112)
# --------------------------------------------------------------------------------------
_SOBEL = torch.tensor([[[[-1., 0., 1.], [-2., 0., 2.], [-1., 0., 1.]]],
                       [[[-1., -2., -1.], [0., 0., 0.], [1., 2., 1.]]]])  # shi_tomasi.py:45-59


def shi_tomasi_score(image: torch.Tensor, block_size: int = 3, ieee_sqrt: bool = False) -> torch.Tensor:
    """ieee_sqrt=False is the reference verbatim.  NOTE: torch.sqrt on CPU goes through MKL VML
    (vsSqrt, HA mode), which is NOT correctly rounded: it differs from IEEE-754 sqrt by 1 ulp on
    ~0.6 % of inputs, and the pattern depends on the MKL code path of the host CPU.  The reference's
    score map is therefore not reproducible to the last bit even between two x86 hosts.
    ieee_sqrt=True swaps in the correctly rounded sqrt (numpy) -- the variant a CUDA kernel using
    sqrt.rn.f32 must match bit-for-bit on integer-valued images."""
    img = image.float()                                                   # :78
    g = F.conv2d(F.pad(img, (1, 1, 1, 1), mode="replicate"), _SOBEL)      # :82-83 (cross-correlation)
    ix, iy = g[:, 0:1], g[:, 1:2]
    prod = torch.cat([ix * ix, iy * iy, ix * iy], dim=1)                  # :88
    b = block_size // 2
    ones = torch.ones(3, 1, block_size, block_size)
    s = F.conv2d(F.pad(prod, (b, b, b, b), mode="replicate"), ones, groups=3)   # :92-93 (sum, not mean)
    a, c, bb = s[:, 0:1], s[:, 1:2], s[:, 2:3]
    half_trace = (a + c) / 2                                              # :102
    diff_half = (a - c) / 2                                               # :103
    disc = diff_half * diff_half + bb * bb                                # :104
    arg = disc + 1e-10                                                    # :105
    root = torch.from_numpy(np.sqrt(arg.numpy())) if ieee_sqrt else torch.sqrt(arg)
    lam = half_trace - root                                               # :105-107
    return torch.clamp(lam, min=0.0)                                      # :110


# --------------------------------------------------------------------------------------
# a2/a3: NMS mask and top-k (utils/keypoint_utils.py:12-44, 47-117)
# --------------------------------------------------------------------------------------
def nms_mask(scores: torch.Tensor, nms_radius: int) -> torch.Tensor:
    r = nms_radius
    padded = F.pad(scores.unsqueeze(1), (r, r, r, r), mode="constant", value=float("-inf"))   # :29-34
    local_max = F.max_pool2d(padded, kernel_size=2 * r + 1, stride=1, padding=0).squeeze(1)   # :36-41
    return (scores >= (local_max - 1e-7)).float()                                             # :43


def masked_scores(scores: torch.Tensor, mask: torch.Tensor, score_threshold: float = 0.0,
                  border_margin: int = 0) -> torch.Tensor:
    """The map torch.topk runs over (keypoint_utils.py:77-92)."""
    B, H, W = scores.shape
    if border_margin > 0:
        m = border_margin
        yv = ((torch.arange(H) >= m) & (torch.arange(H) < H - m)).float()
        xv = ((torch.arange(W) >= m) & (torch.arange(W) < W - m)).float()
        sm = scores * mask * (yv.view(1, H, 1) * xv.view(1, 1, W))
    else:
        sm = scores * mask
    return torch.where(sm > score_threshold, sm, torch.zeros_like(sm))


def select_topk(scores: torch.Tensor, mask: torch.Tensor, max_keypoints: int,
                score_threshold: float = 0.0, border_margin: int = 0):
    B, H, W = scores.shape
    sm = masked_scores(scores, mask, score_threshold, border_margin)
    top_s, top_i = torch.topk(sm.reshape(B, -1), k=max_keypoints, dim=1, largest=True, sorted=True)  # :96-102
    kp = torch.stack([(top_i // W).float(), (top_i % W).float()], dim=-1)                           # :104-106
    valid = (top_s > 0).float()                                                                     # :108
    kp = torch.where(valid.unsqueeze(-1) > 0.5, kp, torch.full_like(kp, -1.0))                      # :109-114
    return kp, top_s * valid                                                                        # :115


# --------------------------------------------------------------------------------------
# a8: orientation map (orientation/angle_estimation.py:100-121, 161-170)
# --------------------------------------------------------------------------------------
def moment_kernels(patch_size: int = 15, sigma: float = 2.5) -> torch.Tensor:
    h = patch_size // 2
    y, x = torch.meshgrid(torch.arange(-h, h + 1, dtype=torch.float32),
                          torch.arange(-h, h + 1, dtype=torch.float32), indexing="ij")
    g = torch.exp(-(x ** 2 + y ** 2) / (2 * sigma ** 2))
    return torch.cat([(x * g).view(1, 1, patch_size, patch_size),
                      (y * g).view(1, 1, patch_size, patch_size)], dim=0)


def angle_map(image: torch.Tensor, patch_size: int = 15, sigma: float = 2.5) -> torch.Tensor:
    m = F.conv2d(image, moment_kernels(patch_size, sigma), padding=patch_size // 2)   # zero padding, :161
    return torch.atan2(m[:, 1:2], m[:, 0:1])                                         # :170


# --------------------------------------------------------------------------------------
# a6: sparse BAD (descriptor/bad.py:436-576)
# --------------------------------------------------------------------------------------
def _finish_descriptor(centered, valid, binarize, soft_binarize, temperature, normalize):
    if not binarize:
        d = centered
    elif soft_binarize:
        d = torch.sigmoid(-centered * temperature)
    else:
        d = (centered <= 0).to(centered.dtype)
    d = d * valid.unsqueeze(-1)
    if normalize:
        d = F.normalize(d, p=2, dim=-1)
    return d


def sparse_bad(image: torch.Tensor, keypoints: torch.Tensor, orientation: torch.Tensor | None = None,
               num_pairs: int = 256, binarize: bool = False, soft_binarize: bool = True,
               temperature: float = 10.0, normalize_descriptors: bool = True,
               sampling_mode: str = "nearest") -> torch.Tensor:
    ox1, ox2, oy1, oy2, radii, thr = bad_tables(num_pairs)
    R = int(radii.max())
    _, _, H, W = image.shape
    valid = (keypoints[:, :, 0] >= 0).float()                                  # :461
    yc = torch.clamp(keypoints[:, :, 0], min=0.0, max=float(H - 1))            # :464-465
    xc = torch.clamp(keypoints[:, :, 1], min=0.0, max=float(W - 1))
    sy = 2.0 / (H - 1 + 1e-8)                                                  # :469-470
    sx = 2.0 / (W - 1 + 1e-8)
    bank = F.conv2d(F.pad(image, (R, R, R, R), mode="replicate"), box_kernel_bank(R).to(image.dtype))  # :474-479
    vy1, vx1, vy2, vx2 = (t.view(1, 1, -1) for t in (oy1, ox1, oy2, ox2))
    if orientation is not None:                                                # :487-517
        og = torch.stack([xc * sx - 1.0, yc * sy - 1.0], dim=-1).unsqueeze(2)
        theta = F.grid_sample(orientation, og, mode="nearest", padding_mode="border",
                              align_corners=True).squeeze(1).squeeze(-1)
        ct = torch.cos(theta).unsqueeze(-1)
        st = torch.sin(theta).unsqueeze(-1)
        dy1 = vx1 * st + vy1 * ct
        dx1 = vx1 * ct - vy1 * st
        dy2 = vx2 * st + vy2 * ct
        dx2 = vx2 * ct - vy2 * st
        ky, kx = yc.unsqueeze(-1), xc.unsqueeze(-1)
        p1y, p1x, p2y, p2x = ky + dy1, kx + dx1, ky + dy2, kx + dx2
    else:                                                                      # :518-525
        ky, kx = yc.unsqueeze(-1), xc.unsqueeze(-1)
        p1y, p1x, p2y, p2x = ky + vy1, kx + vx1, ky + vy2, kx + vx2
    g1 = torch.stack([p1x * sx - 1.0, p1y * sy - 1.0], dim=-1)                 # :528-535
    g2 = torch.stack([p2x * sx - 1.0, p2y * sy - 1.0], dim=-1)
    s1 = F.grid_sample(bank, g1, mode=sampling_mode, padding_mode="border", align_corners=True)  # :538-551
    s2 = F.grid_sample(bank, g2, mode=sampling_mode, padding_mode="border", align_corners=True)
    sel = torch.zeros(R + 1, num_pairs)
    sel[radii, torch.arange(num_pairs)] = 1.0                                  # :420-423
    sel = sel.view(1, -1, 1, num_pairs)
    diff = (s1 * sel).sum(dim=1) - (s2 * sel).sum(dim=1)                       # :554-557
    centered = diff - thr.view(1, 1, -1)                                       # :559
    return _finish_descriptor(centered, valid, binarize, soft_binarize, temperature, normalize_descriptors)


# --------------------------------------------------------------------------------------
# a4/a5: dense BAD map and bilinear gather (descriptor/bad.py:62-110, 189-218, 277-333)
# --------------------------------------------------------------------------------------
def dense_bad_diff(x: torch.Tensor, num_pairs: int = 256, pair_chunk: int = 32) -> torch.Tensor:
    """sample1 - sample2 of bad.py:62-110, computed pair-chunk by pair-chunk to bound memory."""
    ox1, ox2, oy1, oy2, radii, _ = bad_tables(num_pairs)
    R = int(radii.max())
    B, _, H, W = x.shape
    xp = F.pad(x, (R, R, R, R), mode="replicate")                               # :70
    integral = torch.cumsum(torch.cumsum(xp, dim=2), dim=3)                     # :71 (double accumulators)
    integral = F.pad(integral, (1, 0, 1, 0), mode="constant", value=0.0).squeeze(1)   # :72
    Wp1 = integral.shape[2]
    flat = integral.reshape(B, -1)
    by = torch.arange(H, dtype=x.dtype).view(1, H, 1)
    bx = torch.arange(W, dtype=x.dtype).view(1, 1, W)
    out = torch.empty(B, num_pairs, H, W, dtype=x.dtype)

    def box_mean(oy, ox, rr):
        cy = torch.clamp(by + oy.view(-1, 1, 1), min=0.0, max=float(H - 1)).to(torch.int64) + R   # :81-85
        cx = torch.clamp(bx + ox.view(-1, 1, 1), min=0.0, max=float(W - 1)).to(torch.int64) + R
        r = rr.view(-1, 1, 1)
        y0, x0, y1, x1 = cy - r, cx - r, cy + r + 1, cx + r + 1

        def g(yi, xi):
            return flat[:, (yi * Wp1 + xi).reshape(-1)].reshape(B, -1, H, W)

        area_sum = g(y1, x1) - g(y0, x1) - g(y1, x0) + g(y0, x0)                 # :98 (this order)
        return area_sum / ((2.0 * rr.float() + 1.0) ** 2).view(-1, 1, 1)         # :99

    for p0 in range(0, num_pairs, pair_chunk):
        s = slice(p0, min(p0 + pair_chunk, num_pairs))
        out[:, s] = box_mean(oy1[s], ox1[s], radii[s]) - box_mean(oy2[s], ox2[s], radii[s])   # :101-110
    return out


def dense_bad(x: torch.Tensor, num_pairs: int = 256, binarize: bool = False, soft_binarize: bool = True,
              temperature: float = 10.0) -> torch.Tensor:
    _, _, _, _, _, thr = bad_tables(num_pairs)
    centered = dense_bad_diff(x, num_pairs) - thr.view(1, -1, 1, 1)             # :212
    if not binarize:
        return centered
    if soft_binarize:
        return torch.sigmoid(-centered * temperature)                          # :217
    return (centered <= 0).to(centered.dtype)                                  # :218


def gather_subpixel(descriptor_map: torch.Tensor, keypoints: torch.Tensor) -> torch.Tensor:
    """extract_descriptors_at_keypoints_subpixel, bad.py:277-333."""
    B, D, H, W = descriptor_map.shape
    yn = keypoints[:, :, 0] / (H - 1 + 1e-8) * 2.0 - 1.0                        # :311-312
    xn = keypoints[:, :, 1] / (W - 1 + 1e-8) * 2.0 - 1.0
    grid = torch.stack([xn, yn], dim=-1).unsqueeze(2)
    s = F.grid_sample(descriptor_map, grid, mode="bilinear", padding_mode="border", align_corners=True)  # :322-328
    return s.squeeze(-1).permute(0, 2, 1)


def dense_descriptors_at_keypoints(descriptor_map, keypoints, normalize=True):
    """feature_detection/shi_tomasi_bad_sinkhorn.py:120-160 and :212-214."""
    _, _, H, W = descriptor_map.shape
    valid = (keypoints[:, :, 0] >= 0).float()
    kc = torch.stack([torch.clamp(keypoints[:, :, 0], min=0.0, max=float(H - 1)),
                      torch.clamp(keypoints[:, :, 1], min=0.0, max=float(W - 1))], dim=-1)
    d = gather_subpixel(descriptor_map, kc) * valid.unsqueeze(-1)
    return F.normalize(d, p=2, dim=-1) if normalize else d


# --------------------------------------------------------------------------------------
# a9: Sinkhorn (matching/sinkhorn.py:79-208)
# --------------------------------------------------------------------------------------
def cost_matrix(d1: torch.Tensor, d2: torch.Tensor, distance_type: str = "l2") -> torch.Tensor:
    if distance_type == "l2":
        n1 = (d1 ** 2).sum(dim=-1, keepdim=True)
        n2 = (d2 ** 2).sum(dim=-1, keepdim=True)
        c = n1 + n2.transpose(-2, -1) - 2.0 * torch.bmm(d1, d2.transpose(-2, -1))   # :98-101
        return torch.clamp(c, min=0.0)                                              # :103
    return torch.abs(d1.unsqueeze(2) - d2.unsqueeze(1)).sum(dim=-1)                 # :106-108


def sinkhorn(d1: torch.Tensor, d2: torch.Tensor, iterations: int = 20, epsilon: float = 1.0,
             unused_score: float = 1.0, distance_type: str = "l2") -> torch.Tensor:
    B, N, _ = d1.shape
    M = d2.shape[1]
    s = F.pad(-cost_matrix(d1, d2, distance_type) / epsilon, (0, 1, 0, 1), value=-unused_score / epsilon)  # :178-187
    log_m = torch.log(torch.tensor(float(M), dtype=d1.dtype))
    log_n = torch.log(torch.tensor(float(N), dtype=d2.dtype))
    log_mu = torch.cat([d1.new_zeros(B, N), log_m.reshape(1, 1).expand(B, -1)], dim=1)   # :197-200
    log_nu = torch.cat([d2.new_zeros(B, M), log_n.reshape(1, 1).expand(B, -1)], dim=1)
    u = torch.zeros_like(log_mu)
    v = torch.zeros_like(log_nu)
    for _ in range(iterations):                                                          # :138-142
        u = log_mu - torch.logsumexp(s + v.unsqueeze(-2), dim=-1)
        v = log_nu - torch.logsumexp(s + u.unsqueeze(-1), dim=-2)
    return torch.exp(s + u.unsqueeze(-1) + v.unsqueeze(-2))                              # :145, :206


# --------------------------------------------------------------------------------------
# a10: unified pipelines (feature_detection/*.py)
# --------------------------------------------------------------------------------------
def filter_rows(P: torch.Tensor, ratio_threshold: float = -1.0, dustbin_margin: float = -1.0):
    """The filtering part of SinkhornMatcherWithFilters.forward, matching/sinkhorn.py:311-465, same ATen ops."""
    B, N, M = P.shape[0], P.shape[1] - 1, P.shape[2] - 1
    valid = torch.ones(B, N, dtype=torch.bool)
    core = P[:, :N, :M]
    if ratio_threshold > 0:                                              # :332-346
        if M >= 2:
            top2 = torch.topk(core, k=2, dim=2, largest=True, sorted=True).values
            best, second = top2[:, :, 0], top2[:, :, 1]
        else:
            best = core[:, :, 0]
            second = torch.zeros_like(best)
        valid = valid & ((best / (second + 1e-8)) >= ratio_threshold)
    if dustbin_margin >= 0:                                              # :366-376
        valid = valid & ((core.max(dim=2).values - P[:, :N, M]) >= dustbin_margin)
    vf = valid.unsqueeze(-1).float()                                     # :448-457
    rows = torch.cat([P[:, :N, :M] * vf, (1.0 - vf) + vf * P[:, :N, M:M + 1]], dim=-1)
    return torch.cat([rows, P[:, N:N + 1, :]], dim=1), valid


def mutual_matches(P: torch.Tensor, keypoints1: torch.Tensor, keypoints2: torch.Tensor, max_matches: int = 100,
                   threshold: float = 0.1):
    """matching/match_extraction.py:46-184 with the same ATen ops in the same order."""
    B, N, M = P.shape[0], keypoints1.shape[1], keypoints2.shape[1]
    core = P[:, :N, :M]                                                  # :72
    max_j_for_i = torch.argmax(core, dim=2)                              # :76
    max_prob_i = torch.max(core, dim=2).values                           # :77
    max_i_for_j = torch.argmax(core, dim=1)                              # :80
    matched_i = torch.gather(max_i_for_j, 1, max_j_for_i)                # :95-99
    is_mutual = matched_i == torch.arange(N).unsqueeze(0).expand(B, -1)  # :102-103
    valid = is_mutual & (max_prob_i >= threshold)                        # :106-109
    s = torch.where(valid, max_prob_i, torch.tensor(-1.0, dtype=P.dtype))            # :117-121
    ss, si = torch.topk(s, k=min(max_matches, N), dim=1, largest=True, sorted=True)  # :124-130
    if N < max_matches:                                                  # :133-142
        ss = torch.cat([ss, torch.zeros(B, max_matches - N, dtype=P.dtype)], dim=1)
        si = torch.cat([si, torch.zeros(B, max_matches - N, dtype=torch.long)], dim=1)
    idx = si.clamp(0, N - 1)
    mk1 = torch.gather(keypoints1, 1, idx.unsqueeze(-1).expand(-1, -1, 2))           # :151-160
    j = torch.gather(max_j_for_i, 1, idx).clamp(0, M - 1)                            # :164-172
    mk2 = torch.gather(keypoints2, 1, j.unsqueeze(-1).expand(-1, -1, 2))             # :174-178
    return mk1, mk2, ss, ss > 0.0                                        # :181-184


# --------------------------------------------------------------------------------------
# essential-matrix head: geometry/essential_matrix_estimator.py,
# feature_detection/shi_tomasi_angle_sparse_bad_sinkhorn_essential_matrix.py:184-271
# --------------------------------------------------------------------------------------
def _det3(M):
    """geometry/essential_matrix_estimator.py:126-130"""
    return (M[0, 0] * (M[1, 1] * M[2, 2] - M[1, 2] * M[2, 1]) - M[0, 1] * (M[1, 0] * M[2, 2] - M[1, 2] * M[2, 0])
            + M[0, 2] * (M[1, 0] * M[2, 1] - M[1, 1] * M[2, 0]))


def _hartley(pts, w):
    """weighted centroid / RMS-distance normalisation, essential_matrix_estimator.py:242-290"""
    w_sum = w.sum() + 1e-8
    c = (w.unsqueeze(-1) * pts).sum(dim=0) / w_sum
    d2 = ((pts - c) ** 2).sum(dim=-1)
    mean_dist = torch.sqrt((w * d2).sum() / w_sum + 1e-8)
    s = math.sqrt(2.0) / (mean_dist + 1e-8)
    z, o = torch.zeros((), dtype=pts.dtype), torch.ones((), dtype=pts.dtype)
    T = torch.stack([torch.stack([s, z, -s * c[0]]), torch.stack([z, s, -s * c[1]]), torch.stack([z, z, o])])
    return T, s, c


def essential_weights(P, valid1=None, valid2=None, top_k=3):
    """Pair weights: validity masking (..._essential_matrix.py:212-219), bidirectional top-k mask AND P > 0.01
    (:221-239; the grid form: essential_matrix_estimator.py:307-330)."""
    N, M = P.shape[0] - 1, P.shape[1] - 1
    core = P[:N, :M]
    if valid1 is not None:
        core = core * valid1.to(core).unsqueeze(1) * valid2.to(core).unsqueeze(0)
    thr_row = torch.topk(core, k=top_k, dim=1).values[:, top_k - 1:top_k]
    thr_col = torch.topk(core, k=top_k, dim=0).values[top_k - 1:top_k, :]
    mask = (core >= thr_row) & (core >= thr_col) & (core > 0.01)
    return core * mask.to(core.dtype)


def essential_matrix(P, pts1_n, pts2_n, valid1=None, valid2=None, top_k=3, n_iter=30, n_iter_manifold=10,
                     dtype=torch.float32):
    """Weighted 8-point essential matrix of ONE pair from its Sinkhorn matrix P (N+1, M+1) and the normalised (x, y)
    points of both images (..._essential_matrix.py:184-271 == essential_matrix_estimator.py:292-392 with grid points).
    dtype=float64 gives the same algorithm without float32 rounding (used to size test tolerances)."""
    P, pts1_n, pts2_n = P.to(dtype), pts1_n.to(dtype), pts2_n.to(dtype)
    N, M = P.shape[0] - 1, P.shape[1] - 1
    w = essential_weights(P, valid1, valid2, top_k)
    T1, s1, c1 = _hartley(pts1_n, w.sum(dim=1))
    T2, s2, c2 = _hartley(pts2_n, w.sum(dim=0))
    f1 = torch.cat([(pts1_n - c1) * s1, torch.ones(N, 1, dtype=dtype)], dim=-1)
    f2 = torch.cat([(pts2_n - c2) * s2, torch.ones(M, 1, dtype=dtype)], dim=-1)
    F1 = (f1.unsqueeze(-1) * f1.unsqueeze(-2)).reshape(N, 9)
    F2 = (f2.unsqueeze(-1) * f2.unsqueeze(-2)).reshape(M, 9)
    M_flat = F1.T @ (w @ F2)                                                     # :249-251
    M_mat = M_flat.reshape(3, 3, 3, 3).permute(0, 2, 1, 3).reshape(9, 9)
    # minimum eigenvector by shifted power iteration, essential_matrix_estimator.py:150-172
    M_s = torch.einsum("ii", M_mat) * torch.eye(9, dtype=dtype) - M_mat
    v = torch.ones(9, dtype=dtype) / 3.0
    for _ in range(n_iter):
        v = M_s @ v
        v = v / (v.norm() + 1e-8)
    E = T2.T @ v.reshape(3, 3) @ T1                                              # :259
    # projection onto singular values (s, s, 0), essential_matrix_estimator.py:174-240
    B = E.T @ E
    lam = torch.einsum("ii", B)
    v1 = torch.ones(3, dtype=dtype) / math.sqrt(3.0)
    for _ in range(n_iter_manifold):
        v1 = B @ v1
        v1 = v1 / (v1.norm() + 1e-8)
    B_s = lam * torch.eye(3, dtype=dtype) - B
    v3 = torch.ones(3, dtype=dtype) / math.sqrt(3.0)
    for _ in range(n_iter_manifold):
        v3 = B_s @ v3
        v3 = v3 / (v3.norm() + 1e-8)
    v2 = torch.linalg.cross(v3, v1)
    v2 = v2 / (v2.norm() + 1e-8)
    V = torch.stack([v1, v2, v3], dim=-1)
    V = V @ torch.diag(torch.stack([torch.ones((), dtype=dtype), torch.ones((), dtype=dtype), torch.sign(_det3(V))]))
    sigma1, sigma2 = (E @ V[:, 0]).norm(), (E @ V[:, 1]).norm()
    s_avg = (sigma1 + sigma2) / 2.0
    u1, u2 = E @ V[:, 0] / (sigma1 + 1e-8), E @ V[:, 1] / (sigma2 + 1e-8)
    U = torch.stack([u1, u2, torch.linalg.cross(u1, u2)], dim=-1)
    U = U @ torch.diag(torch.stack([torch.ones((), dtype=dtype), torch.ones((), dtype=dtype), torch.sign(_det3(U))]))
    return U @ torch.diag(torch.stack([s_avg, s_avg, torch.zeros((), dtype=dtype)])) @ V.T


def normalised_points(keypoints_yx, K_inv):
    """(y, x) pixel keypoints -> (x, y) normalised image coordinates, ..._essential_matrix.py:341-352"""
    xy = torch.stack([keypoints_yx[:, 1], keypoints_yx[:, 0]], dim=-1)
    return (torch.cat([xy, torch.ones(xy.shape[0], 1, dtype=xy.dtype)], dim=-1) @ K_inv.T.to(xy))[:, :2]


def grid_points(n, image_shape, K_inv):
    """essential_matrix_estimator.py:85-105: feature i sits at pixel (i % W, i // W)"""
    H, W = image_shape
    idx = torch.arange(H * W, dtype=torch.float32)
    h = torch.stack([idx % W, idx // W, torch.ones(H * W)], dim=-1)
    return (h @ K_inv.T)[:n, :2]


# False: detect() takes the square root exactly as the reference does on this host (torch.sqrt: MKL VML, host-dependent in
# the last bit, see shi_tomasi_score).  The GPU parity tests switch it on (fixture in tests/test_gpu_parity.py) so that the
# expected keypoints do not depend on the CPU of the GPU box; the CPU tests pin the verbatim form to the goldens.
DETECT_IEEE_SQRT = False


def detect(image, max_keypoints, block_size=3, nms_radius=3, score_threshold=0.0, border_margin=0):
    sc = shi_tomasi_score(image, block_size, ieee_sqrt=DETECT_IEEE_SQRT).squeeze(1)
    return select_topk(sc, nms_mask(sc, nms_radius), max_keypoints, score_threshold, border_margin)


def sparse_matcher(image1, image2, max_keypoints, block_size=3, num_pairs=256, binarize=False,
                   soft_binarize=True, temperature=10.0, sinkhorn_iterations=20, epsilon=1.0,
                   unused_score=1.0, distance_type="l2", nms_radius=3, score_threshold=0.0,
                   normalize_descriptors=True, sampling_mode="nearest", border_margin=None,
                   return_descriptors=False):
    """feature_detection/shi_tomasi_sparse_bad_sinkhorn.py:134-182."""
    margin = 7 if border_margin is None else border_margin          # descriptor.max_radius, :121-124
    k1, _ = detect(image1, max_keypoints, block_size, nms_radius, score_threshold, margin)
    k2, _ = detect(image2, max_keypoints, block_size, nms_radius, score_threshold, margin)
    kw = dict(num_pairs=num_pairs, binarize=binarize, soft_binarize=soft_binarize, temperature=temperature,
              normalize_descriptors=normalize_descriptors, sampling_mode=sampling_mode)
    d1 = sparse_bad(image1, k1, None, **kw)
    d2 = sparse_bad(image2, k2, None, **kw)
    p = sinkhorn(d1, d2, sinkhorn_iterations, epsilon, unused_score, distance_type)
    return (k1, k2, p, d1, d2) if return_descriptors else (k1, k2, p)


def angle_matcher(image1, image2, max_keypoints, block_size=5, patch_size=15, sigma=2.5, num_pairs=256,
                  binarize=False, soft_binarize=True, temperature=10.0, sinkhorn_iterations=20,
                  epsilon=1.0, unused_score=1.0, distance_type="l2", nms_radius=3, score_threshold=0.0,
                  normalize_descriptors=True, sampling_mode="nearest", border_margin=None,
                  return_descriptors=False):
    """feature_detection/shi_tomasi_angle_sparse_bad_sinkhorn.py:132-180."""
    margin = 7 if border_margin is None else border_margin
    k1, _ = detect(image1, max_keypoints, block_size, nms_radius, score_threshold, margin)
    k2, _ = detect(image2, max_keypoints, block_size, nms_radius, score_threshold, margin)
    kw = dict(num_pairs=num_pairs, binarize=binarize, soft_binarize=soft_binarize, temperature=temperature,
              normalize_descriptors=normalize_descriptors, sampling_mode=sampling_mode)
    d1 = sparse_bad(image1, k1, angle_map(image1, patch_size, sigma), **kw)
    d2 = sparse_bad(image2, k2, angle_map(image2, patch_size, sigma), **kw)
    p = sinkhorn(d1, d2, sinkhorn_iterations, epsilon, unused_score, distance_type)
    return (k1, k2, p, d1, d2) if return_descriptors else (k1, k2, p)


def angle_detector(image, max_keypoints, block_size=5, patch_size=15, sigma=2.5, num_pairs=256,
                   binarize=False, soft_binarize=True, temperature=10.0, normalize_descriptors=True,
                   sampling_mode="nearest", nms_radius=3, score_threshold=0.0):
    """feature_detection/shi_tomasi_angle.py:323-356 (no border margin, returns scores)."""
    k, s = detect(image, max_keypoints, block_size, nms_radius, score_threshold, 0)
    d = sparse_bad(image, k, angle_map(image, patch_size, sigma), num_pairs=num_pairs, binarize=binarize,
                   soft_binarize=soft_binarize, temperature=temperature,
                   normalize_descriptors=normalize_descriptors, sampling_mode=sampling_mode)
    return k, s, d


def dense_detector(image, block_size=3, num_pairs=256, binarize=False, soft_binarize=True, temperature=10.0):
    """feature_detection/shi_tomasi_bad.py:72-89."""
    return shi_tomasi_score(image, block_size), dense_bad(image, num_pairs, binarize, soft_binarize, temperature)


def dense_matcher(image1, image2, max_keypoints, block_size=3, num_pairs=256, binarize=False,
                  soft_binarize=True, temperature=10.0, sinkhorn_iterations=20, epsilon=1.0,
                  unused_score=1.0, distance_type="l2", nms_radius=3, score_threshold=0.0,
                  normalize_descriptors=True, return_descriptors=False):
    """feature_detection/shi_tomasi_bad_sinkhorn.py:162-219 (no border margin).  One image at a time
    so the (P,H,W) map (314 MB at 480x640) is never held for a whole batch."""
    outs = []
    for img in (image1, image2):
        ks, ds = [], []
        for b in range(img.shape[0]):
            one = img[b:b + 1]
            k, _ = detect(one, max_keypoints, block_size, nms_radius, score_threshold, 0)
            dmap = dense_bad(one.float(), num_pairs, binarize, soft_binarize, temperature)
            ds.append(dense_descriptors_at_keypoints(dmap, k, normalize_descriptors))
            ks.append(k)
            del dmap
        outs.append((torch.cat(ks), torch.cat(ds)))
    (k1, d1), (k2, d2) = outs
    p = sinkhorn(d1, d2, sinkhorn_iterations, epsilon, unused_score, distance_type)
    return (k1, k2, p, d1, d2) if return_descriptors else (k1, k2, p)


# --------------------------------------------------------------------------------------
# synthetic inputs shared by tests and bench (integer-valued [0,255] so scores are order-exact)
# --------------------------------------------------------------------------------------
# --------------------------------------------------------------------------------------
# f4: the matcher behind another detector (feature_detection/akaze_sparse_bad_sinkhorn.py:148-196) and the callers' ingest
# --------------------------------------------------------------------------------------
def akaze_maps(image, num_scales=3, diffusion_iterations=3, kappa=0.05, threshold=0.001, nms_size=5, patch_size=15, sigma=2.5):
    """detector/akaze.py:331-453 with its parts (:25-134 diffusion, :137-263 Hessian response, :266-328 orientation): the
    same ATen operators in the same order -> (scores, orientations), both (B,1,H,W)."""
    def k3(rows, scale):
        return (torch.tensor(rows, dtype=torch.float32).view(1, 1, 3, 3)) / scale
    sobel = torch.cat([k3([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], 8.0), k3([[-1, -2, -1], [0, 0, 0], [1, 2, 1]], 8.0)], dim=0)
    hess = torch.cat([k3([[1, -2, 1], [2, -4, 2], [1, -2, 1]], 16.0), k3([[1, 2, 1], [-2, -4, -2], [1, 2, 1]], 16.0),
                      k3([[1, 0, -1], [0, 0, 0], [-1, 0, 1]], 4.0)], dim=0)
    mk = moment_kernels(patch_size, sigma)
    cur, ss, oo = image, [], []
    for _ in range(num_scales):
        for _ in range(diffusion_iterations):                                    # akaze.py:110-132
            g = F.conv2d(cur, sobel, padding=1)
            mag = torch.sqrt((g * g).sum(dim=1, keepdim=True) + 1e-8)
            c = 1.0 / (1.0 + (mag / kappa) ** 2)
            div = F.conv2d(c * g, sobel, padding=1, groups=2).sum(dim=1, keepdim=True)
            cur = cur + 0.25 * div
        h = F.conv2d(cur, hess, padding=1)                                       # :196-205
        resp = h[:, 0:1] * h[:, 1:2] - h[:, 2:3] * h[:, 2:3]
        mx = F.max_pool2d(resp, kernel_size=nms_size, stride=1, padding=nms_size // 2)
        ss.append(torch.clamp(resp * ((resp == mx).float() * (resp > threshold).float()), min=0.0))   # :231-263
        m = F.conv2d(cur, mk, padding=patch_size // 2)
        oo.append(torch.atan2(m[:, 1:2], m[:, 0:1]))
    ss, oo = torch.stack(ss, dim=0), torch.stack(oo, dim=0)
    scores = ss.amax(dim=0)                                                      # :436-449
    mask = (ss == scores.unsqueeze(0)).float()
    mask = mask / mask.sum(dim=0, keepdim=True).clamp(min=1.0)
    return scores, (oo * mask).sum(dim=0)


def maps_matcher(image1, image2, scores1, scores2, orient1, orient2, max_keypoints, num_pairs=256, binarize=False,
                 soft_binarize=True, temperature=10.0, sinkhorn_iterations=20, epsilon=1.0, unused_score=1.0,
                 distance_type="l2", nms_radius=3, score_threshold=0.0, normalize_descriptors=True, sampling_mode="nearest",
                 border_margin=None, return_descriptors=False):
    """feature_detection/akaze_sparse_bad_sinkhorn.py:155-196 from the detector's maps on."""
    margin = 7 if border_margin is None else border_margin
    s1, s2 = scores1.squeeze(1), scores2.squeeze(1)
    k1, _ = select_topk(s1, nms_mask(s1, nms_radius), max_keypoints, score_threshold, margin)
    k2, _ = select_topk(s2, nms_mask(s2, nms_radius), max_keypoints, score_threshold, margin)
    kw = dict(num_pairs=num_pairs, binarize=binarize, soft_binarize=soft_binarize, temperature=temperature,
              normalize_descriptors=normalize_descriptors, sampling_mode=sampling_mode)
    d1 = sparse_bad(image1, k1, orient1, **kw)
    d2 = sparse_bad(image2, k2, orient2, **kw)
    p = sinkhorn(d1, d2, sinkhorn_iterations, epsilon, unused_score, distance_type)
    return (k1, k2, p, d1, d2) if return_descriptors else (k1, k2, p)


def bgr_to_gray_u8(frame: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY) on uint8 (sample/visual_odometry.py:83): OpenCV's 15-bit fixed-point luma."""
    b, g, r = (frame[..., i].astype(np.int64) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def _linear_coefficients(n_out: int, n_in: int, clamp_weight: bool):
    """source index and 11-bit weights per output coordinate.  Horizontally OpenCV moves out-of-range positions to the border
    pixel with weight 0 (clamp_weight); vertically it keeps the fractional weight and only clips the two row indices."""
    scale = 1.0 / (n_out / n_in)                                                 # cv::resize: inv_scale in double, then 1 / it
    f = ((np.arange(n_out) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_weight:
        lo, hi = s < 0, s >= n_in - 1
        f[lo], s[lo] = 0, 0
        f[hi], s[hi] = 0, n_in - 1
    w0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64)      # cvRound: half to even
    w1 = np.rint(f * np.float32(2048)).astype(np.int64)
    return s, w0, w1


def resize_linear_u8(gray: np.ndarray, height: int, width: int) -> np.ndarray:
    """cv2.resize(gray, (width, height), interpolation=cv2.INTER_LINEAR) on uint8 (sample/visual_odometry.py:88): 11-bit
    fixed-point weights, 32-bit horizontal pass, the vector path's vertical pass."""
    h, w = gray.shape
    sx, ax0, ax1 = _linear_coefficients(width, w, True)
    sy, ay0, ay1 = _linear_coefficients(height, h, False)
    S = gray.astype(np.int64)
    rows = S[:, sx] * ax0 + S[:, np.minimum(sx + 1, w - 1)] * ax1
    S0, S1 = rows[np.clip(sy, 0, h - 1)], rows[np.clip(sy + 1, 0, h - 1)]
    out = (((ay0[:, None] * (S0 >> 4)) >> 16) + ((ay1[:, None] * (S1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def load_image_from_array(image: np.ndarray, height: int, width: int) -> np.ndarray:
    """sample/visual_odometry.py:65-92: (H,W,3) BGR or (H,W) uint8 -> (1,1,height,width) float32 in [0,255]."""
    gray = bgr_to_gray_u8(image) if image.ndim == 3 else image
    return resize_linear_u8(gray, height, width).astype(np.float32)[np.newaxis, np.newaxis]


def texture_images(batch: int, H: int, W: int, seed: int = 0, shift=(3, 5)):
    """Family T: smooth random texture + fine noise, min-max to [0,255], rounded; image2 = roll(image1)."""
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(batch, 1, max(H // 8, 2), max(W // 8, 2), generator=g)
    x = F.interpolate(low, size=(H, W), mode="bilinear", align_corners=False)
    x = x + 0.15 * torch.rand(batch, 1, H, W, generator=g)
    lo = x.amin(dim=(2, 3), keepdim=True)
    hi = x.amax(dim=(2, 3), keepdim=True)
    img1 = torch.round((x - lo) / (hi - lo) * 255.0)
    img2 = torch.roll(img1, shifts=shift, dims=(2, 3))
    return img1.contiguous(), img2.contiguous()


def noise_images(batch: int, H: int, W: int, seed: int = 0):
    """Family N: i.i.d. integers in [0,255] (adversarial for stencil exactness)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (batch, 1, H, W), generator=g).float()
