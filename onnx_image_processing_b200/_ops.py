"""torch custom ops (namespace ``b200match``) over the C ABI of libom_b200.so.

Each op passes raw device pointers and the current CUDA stream to one ``om_*`` entry point of
include/om_b200.h.  The ops are registered for CUDA only: calling them with CPU tensors raises,
there is no CPU implementation behind them.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import torch

from . import _native as nat

DESC_RAW, DESC_SOFT, DESC_HARD = 0, 1, 2
SAMPLE_NEAREST, SAMPLE_BILINEAR = 0, 1
THETA_NONE, THETA_MAP, THETA_MOMENTS = 0, 1, 2
MATCH_SPARSE, MATCH_ANGLE, MATCH_DENSE, MATCH_MAPS = 0, 1, 2, 3


def desc_mode(binarize: bool, soft_binarize: bool) -> int:
    if not binarize:
        return DESC_RAW
    return DESC_SOFT if soft_binarize else DESC_HARD


def sampling_code(mode: str) -> int:
    return SAMPLE_BILINEAR if mode == "bilinear" else SAMPLE_NEAREST


def _f32(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor: onnx_image_processing_b200 has no CPU path")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _begin(t: torch.Tensor):
    """(library, stream): the current torch stream of the tensor's device.  The C entry points switch to the device that
    owns their first pointer and restore the caller's, so nothing here touches the current device."""
    return nat.lib(), ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ws(nbytes: int, like: torch.Tensor) -> torch.Tensor:
    ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=like.device)
    if os.environ.get("OM_POISON_WS"):       # tests: no kernel may depend on what a workspace held before
        ws.fill_(0xFF)
    return ws


def _p(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _images(image: torch.Tensor, what: str) -> Tuple[torch.Tensor, int, int, int]:
    if image.dim() == 4:
        if image.shape[1] != 1:
            raise RuntimeError(f"{what}: expected (B,1,H,W), got {tuple(image.shape)}")
        B, _, H, W = image.shape
    elif image.dim() == 3:
        B, H, W = image.shape
    else:
        raise RuntimeError(f"{what}: expected (B,1,H,W) or (B,H,W), got {tuple(image.shape)}")
    return _f32(image, what), B, H, W


def _images_native(image: torch.Tensor, what: str) -> Tuple[torch.Tensor, int, int, int, int]:
    """(tensor, B, H, W, image_dtype code): uint8 images stay uint8 (read natively by the fused kernels: an extension for
    8-bit camera frames, exact), everything else is float32 as in the reference (`img.float()`, shi_tomasi.py:78)."""
    if image.is_cuda and image.dtype == torch.uint8:
        if image.dim() == 4 and image.shape[1] != 1:
            raise RuntimeError(f"{what}: expected (B,1,H,W), got {tuple(image.shape)}")
        if image.dim() not in (3, 4):
            raise RuntimeError(f"{what}: expected (B,1,H,W) or (B,H,W), got {tuple(image.shape)}")
        return image.contiguous(), image.shape[0], image.shape[-2], image.shape[-1], 1
    img, B, H, W = _images(image, what)
    return img, B, H, W, 0


# ---------------------------------------------------------------------------------------------
@torch.library.custom_op("b200match::shi_tomasi_score", mutates_args=(), device_types="cuda")
def shi_tomasi_score(image: torch.Tensor, block_size: int) -> torch.Tensor:
    img, B, H, W = _images(image, "image")
    lib, st = _begin(img)
    out = torch.empty((B, 1, H, W), dtype=torch.float32, device=img.device)
    nat.check(lib.om_shi_tomasi_score_f32(_p(img), B, H, W, block_size, _p(out), st), "om_shi_tomasi_score_f32")
    return out


@shi_tomasi_score.register_fake
def _(image, block_size):
    return image.new_empty((image.shape[0], 1, image.shape[-2], image.shape[-1]), dtype=torch.float32)


@torch.library.custom_op("b200match::nms_mask", mutates_args=(), device_types="cuda")
def nms_mask(scores: torch.Tensor, nms_radius: int) -> torch.Tensor:
    sc, B, H, W = _images(scores, "scores")
    lib, st = _begin(sc)
    out = torch.empty((B, H, W), dtype=torch.float32, device=sc.device)
    nat.check(lib.om_nms_mask_f32(_p(sc), B, H, W, nms_radius, _p(out), st), "om_nms_mask_f32")
    return out


@nms_mask.register_fake
def _(scores, nms_radius):
    return scores.new_empty(tuple(scores.shape), dtype=torch.float32)


@torch.library.custom_op("b200match::select_topk", mutates_args=(), device_types="cuda")
def select_topk(scores: torch.Tensor, mask: torch.Tensor, max_keypoints: int, score_threshold: float,
                border_margin: int) -> Tuple[torch.Tensor, torch.Tensor]:
    sc, B, H, W = _images(scores, "scores")
    mk = _f32(mask, "nms_mask")
    if mk.numel() != sc.numel():
        raise RuntimeError("scores and nms_mask must have the same number of elements")
    if max_keypoints > H * W:
        raise RuntimeError("selected index k out of range")   # what torch.topk raises in the reference
    lib, st = _begin(sc)
    K = max_keypoints
    kpts = torch.empty((B, K, 2), dtype=torch.float32, device=sc.device)
    ks = torch.empty((B, K), dtype=torch.float32, device=sc.device)
    nbytes = lib.om_topk_workspace_bytes(B, H, W, K)
    ws = _ws(nbytes, sc)
    nat.check(lib.om_select_topk_f32(_p(sc), _p(mk), B, H, W, K, float(score_threshold), int(border_margin), _p(kpts),
                                     _p(ks), _p(ws), ws.numel(), st), "om_select_topk_f32")
    return kpts, ks


@select_topk.register_fake
def _(scores, mask, max_keypoints, score_threshold, border_margin):
    B = scores.shape[0]
    return (scores.new_empty((B, max_keypoints, 2), dtype=torch.float32),
            scores.new_empty((B, max_keypoints), dtype=torch.float32))


@torch.library.custom_op("b200match::detect", mutates_args=(), device_types="cuda")
def detect(image: torch.Tensor, max_keypoints: int, block_size: int, nms_radius: int, score_threshold: float,
           border_margin: int) -> Tuple[torch.Tensor, torch.Tensor]:
    img, B, H, W, u8 = _images_native(image, "image")
    if max_keypoints > H * W:
        raise RuntimeError("selected index k out of range")
    lib, st = _begin(img)
    K = max_keypoints
    kpts = torch.empty((B, K, 2), dtype=torch.float32, device=img.device)
    ks = torch.empty((B, K), dtype=torch.float32, device=img.device)
    ws = _ws(lib.om_topk_workspace_bytes(B, H, W, K), img)
    fn = lib.om_detect_u8 if u8 else lib.om_detect_f32
    nat.check(fn(_p(img), B, H, W, block_size, nms_radius, int(border_margin), float(score_threshold),
                 K, _p(None), _p(kpts), _p(ks), _p(ws), ws.numel(), st), "om_detect")
    return kpts, ks


@detect.register_fake
def _(image, max_keypoints, block_size, nms_radius, score_threshold, border_margin):
    B = image.shape[0]
    return (image.new_empty((B, max_keypoints, 2), dtype=torch.float32),
            image.new_empty((B, max_keypoints), dtype=torch.float32))


@torch.library.custom_op("b200match::angle_map", mutates_args=(), device_types="cuda")
def angle_map(image: torch.Tensor, moment_kernels: torch.Tensor) -> torch.Tensor:
    img, B, H, W = _images(image, "image")
    mk = _f32(moment_kernels, "moment_kernels")
    ps = int(mk.shape[-1])
    lib, st = _begin(img)
    out = torch.empty((B, 1, H, W), dtype=torch.float32, device=img.device)
    nat.check(lib.om_angle_map_f32(_p(img), B, H, W, _p(mk), ps, _p(out), st), "om_angle_map_f32")
    return out


@angle_map.register_fake
def _(image, moment_kernels):
    return image.new_empty((image.shape[0], 1, image.shape[-2], image.shape[-1]), dtype=torch.float32)


@torch.library.custom_op("b200match::sparse_bad", mutates_args=(), device_types="cuda")
def sparse_bad(image: torch.Tensor, keypoints: torch.Tensor, pair_table: torch.Tensor, mode: int, temperature: float,
               normalize: bool, sampling: int, theta_mode: int, orientation: Optional[torch.Tensor],
               moment_kernels: Optional[torch.Tensor]) -> torch.Tensor:
    img, B, H, W = _images(image, "image")
    kp = _f32(keypoints, "keypoints")
    tb = _f32(pair_table, "pair_table")
    K, P = int(kp.shape[1]), int(tb.shape[0])
    ori = _f32(orientation, "orientation") if orientation is not None else None
    mk = _f32(moment_kernels, "moment_kernels") if moment_kernels is not None else None
    ps = int(mk.shape[-1]) if mk is not None else 0
    lib, st = _begin(img)
    out = torch.empty((B, K, P), dtype=torch.float32, device=img.device)
    ws = _ws(lib.om_sparse_bad_workspace_bytes(B, H, W, theta_mode), img)
    nat.check(lib.om_sparse_bad_f32(_p(img), B, H, W, _p(kp), K, _p(tb), P, mode, float(temperature), int(normalize),
                                    sampling, theta_mode, _p(ori), _p(mk), ps, _p(out), _p(ws), ws.numel(), st),
              "om_sparse_bad_f32")
    return out


@sparse_bad.register_fake
def _(image, keypoints, pair_table, mode, temperature, normalize, sampling, theta_mode, orientation, moment_kernels):
    return image.new_empty((keypoints.shape[0], keypoints.shape[1], pair_table.shape[0]), dtype=torch.float32)


@torch.library.custom_op("b200match::dense_bad", mutates_args=(), device_types="cuda")
def dense_bad(image: torch.Tensor, pair_table: torch.Tensor, mode: int, temperature: float) -> torch.Tensor:
    img, B, H, W = _images(image, "image")
    tb = _f32(pair_table, "pair_table")
    P = int(tb.shape[0])
    lib, st = _begin(img)
    out = torch.empty((B, P, H, W), dtype=torch.float32, device=img.device)
    ws = _ws(lib.om_dense_bad_workspace_bytes(B, H, W), img)
    nat.check(lib.om_dense_bad_f32(_p(img), B, H, W, _p(tb), P, mode, float(temperature), _p(out), _p(ws), ws.numel(),
                                   st), "om_dense_bad_f32")
    return out


@dense_bad.register_fake
def _(image, pair_table, mode, temperature):
    return image.new_empty((image.shape[0], pair_table.shape[0], image.shape[-2], image.shape[-1]), dtype=torch.float32)


@torch.library.custom_op("b200match::dense_bad_at_keypoints", mutates_args=(), device_types="cuda")
def dense_bad_at_keypoints(image: torch.Tensor, keypoints: torch.Tensor, pair_table: torch.Tensor, mode: int,
                           temperature: float, normalize: bool) -> torch.Tensor:
    img, B, H, W = _images(image, "image")
    kp = _f32(keypoints, "keypoints")
    tb = _f32(pair_table, "pair_table")
    K, P = int(kp.shape[1]), int(tb.shape[0])
    lib, st = _begin(img)
    out = torch.empty((B, K, P), dtype=torch.float32, device=img.device)
    ws = _ws(lib.om_dense_bad_workspace_bytes(B, H, W), img)
    nat.check(lib.om_dense_bad_at_kpts_f32(_p(img), B, H, W, _p(kp), K, _p(tb), P, mode, float(temperature),
                                           int(normalize), _p(out), _p(ws), ws.numel(), st),
              "om_dense_bad_at_kpts_f32")
    return out


@dense_bad_at_keypoints.register_fake
def _(image, keypoints, pair_table, mode, temperature, normalize):
    return image.new_empty((keypoints.shape[0], keypoints.shape[1], pair_table.shape[0]), dtype=torch.float32)


@torch.library.custom_op("b200match::gather_descriptors", mutates_args=(), device_types="cuda")
def gather_descriptors(descriptor_map: torch.Tensor, keypoints: torch.Tensor, subpixel: bool) -> torch.Tensor:
    dm = _f32(descriptor_map, "descriptor_map")
    kp = _f32(keypoints, "keypoints")
    B, D, H, W = dm.shape
    K = int(kp.shape[1])
    lib, st = _begin(dm)
    out = torch.empty((B, K, D), dtype=torch.float32, device=dm.device)
    nat.check(lib.om_gather_descriptors_f32(_p(dm), B, D, H, W, _p(kp), K, int(subpixel), _p(out), st),
              "om_gather_descriptors_f32")
    return out


@gather_descriptors.register_fake
def _(descriptor_map, keypoints, subpixel):
    return descriptor_map.new_empty((keypoints.shape[0], keypoints.shape[1], descriptor_map.shape[1]),
                                    dtype=torch.float32)


@torch.library.custom_op("b200match::sinkhorn", mutates_args=(), device_types="cuda")
def sinkhorn(desc1: torch.Tensor, desc2: torch.Tensor, iterations: int, epsilon: float, unused_score: float,
             distance_l1: bool) -> torch.Tensor:
    d1 = _f32(desc1, "desc1")
    d2 = _f32(desc2, "desc2")
    B, N, D = d1.shape
    M = int(d2.shape[1])
    if d2.shape[0] != B or d2.shape[2] != D:
        raise RuntimeError(f"descriptor shapes do not match: {tuple(d1.shape)} vs {tuple(d2.shape)}")
    lib, st = _begin(d1)
    out = torch.empty((B, N + 1, M + 1), dtype=torch.float32, device=d1.device)
    ws = _ws(lib.om_sinkhorn_workspace_bytes(B, N, M, D), d1)
    nat.check(lib.om_sinkhorn_f32(_p(d1), _p(d2), B, N, M, D, iterations, float(epsilon), float(unused_score),
                                  int(distance_l1), _p(out), _p(ws), ws.numel(), st), "om_sinkhorn_f32")
    return out


@sinkhorn.register_fake
def _(desc1, desc2, iterations, epsilon, unused_score, distance_l1):
    return desc1.new_empty((desc1.shape[0], desc1.shape[1] + 1, desc2.shape[1] + 1), dtype=torch.float32)


def _epilogue_outputs(dev, B: int, N: int, M: int, want_probs: bool, want_scores: bool, use_filters: bool, ratio: float,
                      margin: float, max_matches: int, match_threshold: float, kpts1=None, kpts2=None):
    """(struct om_sinkhorn_outputs, tensors): probs (B,N+1,M+1), scores0 (B,N), scores1 (B,M), filter_valid (B,N) uint8,
    matched_kpts1/2 (B,max_matches,2), match_scores (B,max_matches), match_valid (B,max_matches) uint8 -- an unrequested
    piece is an empty tensor and a NULL pointer."""
    f = dict(dtype=torch.float32, device=dev)
    f8 = dict(dtype=torch.uint8, device=dev)
    mm = int(max_matches)
    t = dict(                                      # (every output its own tensor: custom ops may not return aliases)
        probs=torch.empty((B, N + 1, M + 1) if want_probs else (0,), **f),
        scores0=torch.empty((B, N) if want_scores else (0,), **f),
        scores1=torch.empty((B, M) if want_scores else (0,), **f),
        filter_valid=torch.empty((B, N) if use_filters else (0,), **f8),
        mk1=torch.empty((B, mm, 2) if mm > 0 else (0,), **f),
        mk2=torch.empty((B, mm, 2) if mm > 0 else (0,), **f),
        mscores=torch.empty((B, mm) if mm > 0 else (0,), **f),
        mvalid=torch.empty((B, mm) if mm > 0 else (0,), **f8),
    )
    q = lambda x: x.data_ptr() if x.numel() else None  # noqa: E731
    out = nat.SinkhornOutputs(q(t["probs"]), q(t["scores0"]), q(t["scores1"]), int(use_filters), float(ratio), float(margin),
                              q(t["filter_valid"]), int(mm > 0), kpts1.data_ptr() if kpts1 is not None else None,
                              kpts2.data_ptr() if kpts2 is not None else None, mm, float(match_threshold), q(t["mk1"]), q(t["mk2"]),
                              q(t["mscores"]), q(t["mvalid"]))
    return out, t


def _fake_epilogue(like, B, N, M, want_probs, want_scores, use_filters, max_matches):
    f = dict(dtype=torch.float32)
    mm = int(max_matches)
    return (like.new_empty((B, N + 1, M + 1) if want_probs else (0,), **f), like.new_empty((B, N) if want_scores else (0,), **f),
            like.new_empty((B, M) if want_scores else (0,), **f), like.new_empty((B, N) if use_filters else (0,), dtype=torch.bool),
            like.new_empty((B, mm, 2) if mm else (0,), **f), like.new_empty((B, mm, 2) if mm else (0,), **f),
            like.new_empty((B, mm) if mm else (0,), **f), like.new_empty((B, mm) if mm else (0,), dtype=torch.bool))


@torch.library.custom_op("b200match::sinkhorn_ex", mutates_args=(), device_types="cuda")
def sinkhorn_ex(desc1: torch.Tensor, desc2: torch.Tensor, iterations: int, epsilon: float, unused_score: float,
                distance_l1: bool, want_probs: bool, want_scores: bool, use_filters: bool, ratio_threshold: float,
                dustbin_margin: float, keypoints1: Optional[torch.Tensor], keypoints2: Optional[torch.Tensor], max_matches: int,
                match_threshold: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor,
                                                 torch.Tensor, torch.Tensor, torch.Tensor]:
    """Sinkhorn with the matching stage's optional outputs in ONE call (om_sinkhorn_ex_f32): (probs, scores0, scores1,
    filter_valid, matched_kpts1, matched_kpts2, match_scores, match_valid); unrequested pieces are empty tensors.  With the
    tcgen05 cluster kernel everything is computed in its epilogue and probs may be skipped altogether."""
    d1 = _f32(desc1, "desc1")
    d2 = _f32(desc2, "desc2")
    B, N, D = d1.shape
    M = int(d2.shape[1])
    if d2.shape[0] != B or d2.shape[2] != D:
        raise RuntimeError(f"descriptor shapes do not match: {tuple(d1.shape)} vs {tuple(d2.shape)}")
    k1 = _f32(keypoints1, "keypoints1") if max_matches > 0 else None
    k2 = _f32(keypoints2, "keypoints2") if max_matches > 0 else None
    lib, st = _begin(d1)
    out, t = _epilogue_outputs(d1.device, B, N, M, want_probs, want_scores, use_filters, ratio_threshold, dustbin_margin,
                               max_matches, match_threshold, k1, k2)
    ws = _ws(lib.om_sinkhorn_ex_workspace_bytes(B, N, M, D), d1)
    nat.check(lib.om_sinkhorn_ex_f32(_p(d1), _p(d2), B, N, M, D, iterations, float(epsilon), float(unused_score),
                                     int(distance_l1), ctypes.byref(out), _p(ws), ws.numel(), st), "om_sinkhorn_ex_f32")
    return (t["probs"], t["scores0"], t["scores1"], t["filter_valid"].to(torch.bool), t["mk1"], t["mk2"], t["mscores"],
            t["mvalid"].to(torch.bool))


@sinkhorn_ex.register_fake
def _(desc1, desc2, iterations, epsilon, unused_score, distance_l1, want_probs, want_scores, use_filters, ratio_threshold,
      dustbin_margin, keypoints1, keypoints2, max_matches, match_threshold):
    return _fake_epilogue(desc1, desc1.shape[0], desc1.shape[1], desc2.shape[1], want_probs, want_scores, use_filters, max_matches)


@torch.library.custom_op("b200match::filter_rows", mutates_args=(), device_types="cuda")
def filter_rows(probs: torch.Tensor, ratio_threshold: float, dustbin_margin: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """(filtered copy of P, valid mask): the outlier filters of SinkhornMatcherWithFilters."""
    p = _f32(probs, "P").clone()            # the op must not mutate its argument; matchers that own P use filter_rows_
    B, N, M = int(p.shape[0]), int(p.shape[1]) - 1, int(p.shape[2]) - 1
    lib, st = _begin(p)
    valid = torch.empty((B, N), dtype=torch.uint8, device=p.device)
    nat.check(lib.om_sinkhorn_filter_rows_f32(_p(p), B, N, M, float(ratio_threshold), float(dustbin_margin), _p(valid), st),
              "om_sinkhorn_filter_rows_f32")
    return p, valid.to(torch.bool)


@filter_rows.register_fake
def _(probs, ratio_threshold, dustbin_margin):
    return probs.new_empty(tuple(probs.shape)), probs.new_empty((probs.shape[0], probs.shape[1] - 1), dtype=torch.bool)


@torch.library.custom_op("b200match::filter_rows_", mutates_args=("probs",), device_types="cuda")
def filter_rows_(probs: torch.Tensor, ratio_threshold: float, dustbin_margin: float) -> torch.Tensor:
    """In-place form for the matchers that own P (no copy of the (K+1)^2 matrix): filters `probs`, returns the mask."""
    if not (probs.is_cuda and probs.dtype == torch.float32 and probs.is_contiguous()):
        raise RuntimeError("filter_rows_ needs a contiguous float32 CUDA tensor")
    B, N, M = int(probs.shape[0]), int(probs.shape[1]) - 1, int(probs.shape[2]) - 1
    lib, st = _begin(probs)
    valid = torch.empty((B, N), dtype=torch.uint8, device=probs.device)
    nat.check(lib.om_sinkhorn_filter_rows_f32(_p(probs), B, N, M, float(ratio_threshold), float(dustbin_margin), _p(valid), st),
              "om_sinkhorn_filter_rows_f32")
    return valid.to(torch.bool)


@filter_rows_.register_fake
def _(probs, ratio_threshold, dustbin_margin):
    return probs.new_empty((probs.shape[0], probs.shape[1] - 1), dtype=torch.bool)


@torch.library.custom_op("b200match::sinkhorn_scores", mutates_args=(), device_types="cuda")
def sinkhorn_scores(probs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(scores0 (B,N), scores1 (B,M)) of SinkhornMatcherWithScores: row / column maxima of the core block."""
    p = _f32(probs, "P")
    B, N, M = int(p.shape[0]), int(p.shape[1]) - 1, int(p.shape[2]) - 1
    lib, st = _begin(p)
    s0 = torch.empty((B, N), dtype=torch.float32, device=p.device)
    s1 = torch.empty((B, M), dtype=torch.float32, device=p.device)
    nat.check(lib.om_sinkhorn_scores_f32(_p(p), B, N, M, _p(s0), _p(s1), st), "om_sinkhorn_scores_f32")
    return s0, s1


@sinkhorn_scores.register_fake
def _(probs):
    return (probs.new_empty((probs.shape[0], probs.shape[1] - 1)), probs.new_empty((probs.shape[0], probs.shape[2] - 1)))


@torch.library.custom_op("b200match::mutual_matches", mutates_args=(), device_types="cuda")
def mutual_matches(probs: torch.Tensor, keypoints1: torch.Tensor, keypoints2: torch.Tensor, max_matches: int,
                   threshold: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    p = _f32(probs, "P")
    k1 = _f32(keypoints1, "keypoints1")
    k2 = _f32(keypoints2, "keypoints2")
    B, N, M = int(p.shape[0]), int(k1.shape[1]), int(k2.shape[1])
    if p.shape[1] != N + 1 or p.shape[2] != M + 1:
        raise RuntimeError(f"P must be (B,{N + 1},{M + 1}) for these keypoints, got {tuple(p.shape)}")
    lib, st = _begin(p)
    mk1 = torch.empty((B, max_matches, 2), dtype=torch.float32, device=p.device)
    mk2 = torch.empty((B, max_matches, 2), dtype=torch.float32, device=p.device)
    sc = torch.empty((B, max_matches), dtype=torch.float32, device=p.device)
    valid = torch.empty((B, max_matches), dtype=torch.uint8, device=p.device)
    ws = _ws(lib.om_mutual_matches_workspace_bytes(B, N, M), p)
    nat.check(lib.om_mutual_matches_f32(_p(p), _p(k1), _p(k2), B, N, M, int(max_matches), float(threshold), _p(mk1),
                                        _p(mk2), _p(sc), _p(valid), _p(ws), ws.numel(), st), "om_mutual_matches_f32")
    return mk1, mk2, sc, valid.to(torch.bool)


@mutual_matches.register_fake
def _(probs, keypoints1, keypoints2, max_matches, threshold):
    B = probs.shape[0]
    return (probs.new_empty((B, max_matches, 2), dtype=torch.float32), probs.new_empty((B, max_matches, 2), dtype=torch.float32),
            probs.new_empty((B, max_matches), dtype=torch.float32), probs.new_empty((B, max_matches), dtype=torch.bool))


@torch.library.custom_op("b200match::essential_matrix", mutates_args=(), device_types="cuda")
def essential_matrix(probs: torch.Tensor, pts1: torch.Tensor, pts2: torch.Tensor, valid1: Optional[torch.Tensor],
                     valid2: Optional[torch.Tensor], top_k: int, n_iter: int, n_iter_manifold: int) -> torch.Tensor:
    """probs (B,N+1,M+1); pts1 (B,N,2) or (N,2), pts2 (B,M,2) or (M,2): normalised (x, y); valid1/2 (B,N)/(B,M) bool or
    None -> E (B,3,3)."""
    p = _f32(probs, "P")
    a, b = _f32(pts1, "pts1"), _f32(pts2, "pts2")
    B, N, M = int(p.shape[0]), int(p.shape[1]) - 1, int(p.shape[2]) - 1
    batched = a.dim() == 3
    if (b.dim() == 3) != batched or tuple(a.shape[-2:]) != (N, 2) or tuple(b.shape[-2:]) != (M, 2) or \
            (batched and (a.shape[0] != B or b.shape[0] != B)):
        raise RuntimeError(f"points must be ([B,]{N},2) and ([B,]{M},2) for P {tuple(p.shape)}, got {tuple(a.shape)}, {tuple(b.shape)}")
    if top_k > min(N, M):
        raise RuntimeError("selected index k out of range")              # torch.topk's message
    v1 = v2 = None
    if (valid1 is None) != (valid2 is None):
        raise RuntimeError("valid1 and valid2 must be given together (both masks or neither)")
    if valid1 is not None:
        v1 = valid1.to(device=p.device, dtype=torch.uint8).contiguous()
        v2 = valid2.to(device=p.device, dtype=torch.uint8).contiguous()
        if tuple(v1.shape) != (B, N) or tuple(v2.shape) != (B, M):
            raise RuntimeError(f"valid masks must be ({B},{N}) and ({B},{M}), got {tuple(v1.shape)}, {tuple(v2.shape)}")
    lib, st = _begin(p)
    E = torch.empty((B, 3, 3), dtype=torch.float32, device=p.device)
    nat.check(lib.om_essential_matrix_f32(_p(p), _p(a), _p(b), _p(v1), _p(v2), B, N, M, int(batched), int(top_k), int(n_iter),
                                          int(n_iter_manifold), _p(E), st), "om_essential_matrix_f32")
    return E


@essential_matrix.register_fake
def _(probs, pts1, pts2, valid1, valid2, top_k, n_iter, n_iter_manifold):
    return probs.new_empty((probs.shape[0], 3, 3), dtype=torch.float32)


@torch.library.custom_op("b200match::detect_from_scores", mutates_args=(), device_types="cuda")
def detect_from_scores(scores: torch.Tensor, max_keypoints: int, nms_radius: int, score_threshold: float,
                       border_margin: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """apply_nms_maxpool + select_topk_keypoints on a caller's score map in one call, no mask array
    (feature_detection/akaze_sparse_bad_sinkhorn.py:155-170)."""
    sc, B, H, W = _images(scores, "scores")
    if max_keypoints > H * W:
        raise RuntimeError("selected index k out of range")
    lib, st = _begin(sc)
    K = max_keypoints
    kpts = torch.empty((B, K, 2), dtype=torch.float32, device=sc.device)
    ks = torch.empty((B, K), dtype=torch.float32, device=sc.device)
    ws = _ws(lib.om_topk_workspace_bytes(B, H, W, K), sc)
    nat.check(lib.om_detect_from_scores_f32(_p(sc), B, H, W, nms_radius, int(border_margin), float(score_threshold), K,
                                            _p(kpts), _p(ks), _p(ws), ws.numel(), st), "om_detect_from_scores_f32")
    return kpts, ks


@detect_from_scores.register_fake
def _(scores, max_keypoints, nms_radius, score_threshold, border_margin):
    B = scores.shape[0]
    return (scores.new_empty((B, max_keypoints, 2), dtype=torch.float32),
            scores.new_empty((B, max_keypoints), dtype=torch.float32))


@torch.library.custom_op("b200match::preprocess_u8", mutates_args=(), device_types="cuda")
def preprocess_u8(frames: torch.Tensor, height: int, width: int, as_float: bool) -> torch.Tensor:
    """Camera frames (B,Hin,Win,3) BGR or (B,Hin,Win[,1]) grey, uint8 -> (B,1,height,width) grey model input, uint8 or
    float32 (sample/visual_odometry.py:65-92: cv2.cvtColor + cv2.resize(INTER_LINEAR) + astype(float32)), one kernel."""
    if not frames.is_cuda or frames.dtype != torch.uint8:
        raise RuntimeError("frames must be a CUDA uint8 tensor: onnx_image_processing_b200 has no CPU path")
    if frames.dim() == 3:
        frames = frames.unsqueeze(-1)
    if frames.dim() != 4 or frames.shape[-1] not in (1, 3):
        raise RuntimeError(f"frames: expected (B,H,W,3), (B,H,W,1) or (B,H,W), got {tuple(frames.shape)}")
    f = frames.contiguous()
    B, Hin, Win, C = f.shape
    lib, st = _begin(f)
    out = torch.empty((B, 1, height, width), dtype=torch.float32 if as_float else torch.uint8, device=f.device)
    nat.check(lib.om_preprocess_u8(_p(f), B, Hin, Win, C, height, width, _p(None if as_float else out),
                                   _p(out if as_float else None), st), "om_preprocess_u8")
    return out


@preprocess_u8.register_fake
def _(frames, height, width, as_float):
    return frames.new_empty((frames.shape[0], 1, height, width), dtype=torch.float32 if as_float else torch.uint8)


@torch.library.custom_op("b200match::match_pairs_from_maps", mutates_args=(), device_types="cuda")
def match_pairs_from_maps(image1: torch.Tensor, image2: torch.Tensor, scores1: torch.Tensor, scores2: torch.Tensor,
                          orient1: Optional[torch.Tensor], orient2: Optional[torch.Tensor], pair_table: torch.Tensor,
                          max_keypoints: int, nms_radius: int, border_margin: int, score_threshold: float, mode: int,
                          temperature: float, normalize: bool, sampling: int, iterations: int, epsilon: float,
                          unused_score: float, distance_l1: bool
                          ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """The matcher behind another detector (om_match_pairs_from_maps_f32): score / orientation maps in, (kpts1, kpts2, probs,
    desc1, desc2) out -- NMS, top-k, oriented sparse BAD and Sinkhorn in one C call."""
    i1, B, H, W = _images(image1, "image1")
    i2, B2, H2, W2 = _images(image2, "image2")
    s1, Bs, Hs, Ws = _images(scores1, "scores1")
    s2, Bt, Ht, Wt = _images(scores2, "scores2")
    if not ((B, H, W) == (B2, H2, W2) == (Bs, Hs, Ws) == (Bt, Ht, Wt)):
        raise RuntimeError("images and score maps must have the same shape")
    if (orient1 is None) != (orient2 is None):
        raise RuntimeError("orientation maps: give both or neither")
    o1 = _images(orient1, "orient1")[0] if orient1 is not None else None
    o2 = _images(orient2, "orient2")[0] if orient2 is not None else None
    K = max_keypoints
    if K > H * W:
        raise RuntimeError("selected index k out of range")
    tb = _f32(pair_table, "pair_table")
    P = int(tb.shape[0])
    lib, st = _begin(i1)
    prm = make_match_params(MATCH_MAPS, B, H, W, K, 1, nms_radius, border_margin, score_threshold, P, mode, temperature,
                            normalize, sampling, 0, iterations, epsilon, unused_score, distance_l1, 0)
    dev = i1.device
    k1 = torch.empty((B, K, 2), dtype=torch.float32, device=dev)
    k2 = torch.empty((B, K, 2), dtype=torch.float32, device=dev)
    probs = torch.empty((B, K + 1, K + 1), dtype=torch.float32, device=dev)
    d1 = torch.empty((B, K, P), dtype=torch.float32, device=dev)
    d2 = torch.empty((B, K, P), dtype=torch.float32, device=dev)
    nbytes = lib.om_match_workspace_bytes(ctypes.byref(prm))
    if nbytes == 0:
        raise RuntimeError("om_match_workspace_bytes rejected the parameters")
    ws = _ws(nbytes, i1)
    nat.check(lib.om_match_pairs_from_maps_f32(ctypes.byref(prm), _p(i1), _p(i2), _p(s1), _p(s2), _p(o1), _p(o2), _p(tb),
                                               _p(k1), _p(k2), _p(probs), _p(d1), _p(d2), _p(ws), ws.numel(), st),
              "om_match_pairs_from_maps_f32")
    return k1, k2, probs, d1, d2


@match_pairs_from_maps.register_fake
def _(image1, image2, scores1, scores2, orient1, orient2, pair_table, max_keypoints, nms_radius, border_margin,
      score_threshold, mode, temperature, normalize, sampling, iterations, epsilon, unused_score, distance_l1):
    B, K, P = image1.shape[0], max_keypoints, pair_table.shape[0]
    f = dict(dtype=torch.float32)
    return (image1.new_empty((B, K, 2), **f), image1.new_empty((B, K, 2), **f),
            image1.new_empty((B, K + 1, K + 1), **f), image1.new_empty((B, K, P), **f),
            image1.new_empty((B, K, P), **f))


def make_match_params(flavour: int, B: int, H: int, W: int, K: int, block_size: int, nms_radius: int,
                      border_margin: int, score_threshold: float, P: int, mode: int, temperature: float,
                      normalize: bool, sampling: int, patch_size: int, iterations: int, epsilon: float,
                      unused_score: float, distance_l1: bool, image_dtype: int = 0) -> nat.MatchParams:
    return nat.MatchParams(flavour, B, H, W, K, block_size, nms_radius, border_margin, score_threshold, P, mode,
                           temperature, int(normalize), sampling, patch_size, iterations, epsilon, unused_score,
                           int(distance_l1), int(image_dtype))


@torch.library.custom_op("b200match::match_pairs", mutates_args=(), device_types="cuda")
def match_pairs(image1: torch.Tensor, image2: torch.Tensor, pair_table: torch.Tensor,
                moment_kernels: Optional[torch.Tensor], flavour: int, max_keypoints: int, block_size: int,
                nms_radius: int, border_margin: int, score_threshold: float, mode: int, temperature: float,
                normalize: bool, sampling: int, iterations: int, epsilon: float, unused_score: float,
                distance_l1: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Whole matcher forward in one C call: (kpts1, kpts2, probs, desc1, desc2)."""
    i1, B, H, W, u8 = _images_native(image1, "image1")
    i2, B2, H2, W2, u8b = _images_native(image2, "image2")
    if (B, H, W) != (B2, H2, W2):
        raise RuntimeError("image1 and image2 must have the same shape")
    if u8 != u8b:                                  # mixed pixel types: widen the uint8 one
        i1, i2, u8 = i1.float(), i2.float(), 0
    K = max_keypoints
    if K > H * W:
        raise RuntimeError("selected index k out of range")
    tb = _f32(pair_table, "pair_table")
    P = int(tb.shape[0])
    mk = _f32(moment_kernels, "moment_kernels") if moment_kernels is not None else None
    ps = int(mk.shape[-1]) if mk is not None else 0
    lib, st = _begin(i1)
    prm = make_match_params(flavour, B, H, W, K, block_size, nms_radius, border_margin, score_threshold, P, mode,
                            temperature, normalize, sampling, ps, iterations, epsilon, unused_score, distance_l1, u8)
    dev = i1.device
    k1 = torch.empty((B, K, 2), dtype=torch.float32, device=dev)
    k2 = torch.empty((B, K, 2), dtype=torch.float32, device=dev)
    probs = torch.empty((B, K + 1, K + 1), dtype=torch.float32, device=dev)
    d1 = torch.empty((B, K, P), dtype=torch.float32, device=dev)
    d2 = torch.empty((B, K, P), dtype=torch.float32, device=dev)
    nbytes = lib.om_match_workspace_bytes(ctypes.byref(prm))
    if nbytes == 0:
        raise RuntimeError("om_match_workspace_bytes rejected the parameters")
    ws = _ws(nbytes, i1)
    nat.check(lib.om_match_pairs(ctypes.byref(prm), _p(i1), _p(i2), _p(tb), _p(mk), _p(k1), _p(k2), _p(probs),
                                 _p(d1), _p(d2), _p(ws), ws.numel(), st), "om_match_pairs")
    return k1, k2, probs, d1, d2


@torch.library.custom_op("b200match::match_pairs_ex", mutates_args=(), device_types="cuda")
def match_pairs_ex(image1: torch.Tensor, image2: torch.Tensor, pair_table: torch.Tensor,
                   moment_kernels: Optional[torch.Tensor], flavour: int, max_keypoints: int, block_size: int,
                   nms_radius: int, border_margin: int, score_threshold: float, mode: int, temperature: float,
                   normalize: bool, sampling: int, iterations: int, epsilon: float, unused_score: float,
                   distance_l1: bool, want_probs: bool, want_scores: bool, use_filters: bool, ratio_threshold: float,
                   dustbin_margin: float, max_matches: int, match_threshold: float
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor,
                              torch.Tensor, torch.Tensor, torch.Tensor]:
    """Whole matcher forward with the matching stage's optional outputs (om_match_pairs_ex): (kpts1, kpts2, probs, scores0,
    scores1, filter_valid, matched_kpts1, matched_kpts2, match_scores, match_valid).  With want_probs=False and K <= 512 the
    (K+1)^2 matrix is never written: matches come out of the Sinkhorn kernel's epilogue."""
    i1, B, H, W, u8 = _images_native(image1, "image1")
    i2, B2, H2, W2, u8b = _images_native(image2, "image2")
    if (B, H, W) != (B2, H2, W2):
        raise RuntimeError("image1 and image2 must have the same shape")
    if u8 != u8b:
        i1, i2, u8 = i1.float(), i2.float(), 0
    K = max_keypoints
    if K > H * W:
        raise RuntimeError("selected index k out of range")
    tb = _f32(pair_table, "pair_table")
    P = int(tb.shape[0])
    mk = _f32(moment_kernels, "moment_kernels") if moment_kernels is not None else None
    ps = int(mk.shape[-1]) if mk is not None else 0
    lib, st = _begin(i1)
    prm = make_match_params(flavour, B, H, W, K, block_size, nms_radius, border_margin, score_threshold, P, mode,
                            temperature, normalize, sampling, ps, iterations, epsilon, unused_score, distance_l1, u8)
    dev = i1.device
    k1 = torch.empty((B, K, 2), dtype=torch.float32, device=dev)
    k2 = torch.empty((B, K, 2), dtype=torch.float32, device=dev)
    out, t = _epilogue_outputs(dev, B, K, K, want_probs, want_scores, use_filters, ratio_threshold, dustbin_margin, max_matches,
                               match_threshold)
    nbytes = lib.om_match_ex_workspace_bytes(ctypes.byref(prm))
    if nbytes == 0:
        raise RuntimeError("om_match_ex_workspace_bytes rejected the parameters")
    ws = _ws(nbytes, i1)
    nat.check(lib.om_match_pairs_ex(ctypes.byref(prm), _p(i1), _p(i2), _p(tb), _p(mk), _p(k1), _p(k2), _p(None), _p(None),
                                    ctypes.byref(out), _p(ws), ws.numel(), st), "om_match_pairs_ex")
    return (k1, k2, t["probs"], t["scores0"], t["scores1"], t["filter_valid"].to(torch.bool), t["mk1"], t["mk2"], t["mscores"],
            t["mvalid"].to(torch.bool))


@match_pairs_ex.register_fake
def _(image1, image2, pair_table, moment_kernels, flavour, max_keypoints, block_size, nms_radius, border_margin,
      score_threshold, mode, temperature, normalize, sampling, iterations, epsilon, unused_score, distance_l1, want_probs,
      want_scores, use_filters, ratio_threshold, dustbin_margin, max_matches, match_threshold):
    B, K = image1.shape[0], max_keypoints
    f = dict(dtype=torch.float32)
    return (image1.new_empty((B, K, 2), **f), image1.new_empty((B, K, 2), **f)) + \
        _fake_epilogue(image1, B, K, K, want_probs, want_scores, use_filters, max_matches)


@match_pairs.register_fake
def _(image1, image2, pair_table, moment_kernels, flavour, max_keypoints, block_size, nms_radius, border_margin,
      score_threshold, mode, temperature, normalize, sampling, iterations, epsilon, unused_score, distance_l1):
    B, K, P = image1.shape[0], max_keypoints, pair_table.shape[0]
    f = dict(dtype=torch.float32)
    return (image1.new_empty((B, K, 2), **f), image1.new_empty((B, K, 2), **f),
            image1.new_empty((B, K + 1, K + 1), **f), image1.new_empty((B, K, P), **f),
            image1.new_empty((B, K, P), **f))
