"""ctypes binding of libom_b200.so (the C ABI declared in include/om_b200.h).

There is deliberately no fallback: if the library is missing and cannot be built, or a call
fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_size_t, c_ulonglong, c_void_p

from . import build as _build

_lock = threading.Lock()
_lib = None


class MatchParams(Structure):
    """struct om_match_params (include/om_b200.h)."""
    _fields_ = [
        ("flavour", c_int), ("B", c_int), ("H", c_int), ("W", c_int), ("K", c_int),
        ("block_size", c_int), ("nms_radius", c_int), ("border_margin", c_int), ("score_threshold", c_float),
        ("P", c_int), ("desc_mode", c_int), ("temperature", c_float), ("normalize", c_int),
        ("sampling_mode", c_int), ("patch_size", c_int), ("iterations", c_int), ("epsilon", c_float),
        ("unused_score", c_float), ("distance_l1", c_int), ("image_dtype", c_int),
    ]


class SinkhornOutputs(Structure):
    """struct om_sinkhorn_outputs (include/om_b200.h)."""
    _fields_ = [
        ("probs", c_void_p), ("scores0", c_void_p), ("scores1", c_void_p),
        ("filters", c_int), ("ratio_threshold", c_float), ("dustbin_margin", c_float), ("filter_valid", c_void_p),
        ("matches", c_int), ("kpts1", c_void_p), ("kpts2", c_void_p), ("max_matches", c_int), ("match_threshold", c_float),
        ("matched_kpts1", c_void_p), ("matched_kpts2", c_void_p), ("match_scores", c_void_p), ("match_valid", c_void_p),
    ]


# name -> (restype, argtypes); must list every symbol include/om_b200.h declares
SIGNATURES = {
    "om_version": (c_int, []),
    "om_set_device": (c_int, [c_int]),
    "om_error_string": (c_char_p, [c_int]),
    "om_launch_count": (c_ulonglong, []),
    "om_bad_table": (c_int, [c_int, c_void_p, c_void_p]),
    "om_shi_tomasi_score_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "om_nms_mask_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "om_topk_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "om_select_topk_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p,
                                   c_void_p, c_size_t, c_void_p]),
    "om_detect_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_size_t, c_void_p]),
    "om_detect_from_scores_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p,
                                          c_void_p, c_size_t, c_void_p]),
    "om_preprocess_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "om_match_pairs_from_maps_f32": (c_int, [POINTER(MatchParams), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                             c_void_p]),
    "om_angle_map_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "om_sparse_bad_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "om_sparse_bad_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_float,
                                  c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t,
                                  c_void_p]),
    "om_dense_bad_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "om_dense_bad_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p,
                                 c_size_t, c_void_p]),
    "om_dense_bad_at_kpts_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int,
                                         c_float, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "om_gather_descriptors_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p,
                                          c_void_p]),
    "om_sinkhorn_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "om_sinkhorn_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_int,
                                c_void_p, c_void_p, c_size_t, c_void_p]),
    "om_sinkhorn_filter_rows_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p]),
    "om_sinkhorn_scores_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "om_essential_matrix_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                        c_int, c_int, c_void_p, c_void_p]),
    "om_mutual_matches_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "om_mutual_matches_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "om_match_workspace_bytes": (c_size_t, [POINTER(MatchParams)]),
    "om_match_pairs_f32": (c_int, [POINTER(MatchParams), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "om_match_pairs": (c_int, [POINTER(MatchParams), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "om_detect_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_size_t, c_void_p]),
    "om_sinkhorn_ex_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "om_sinkhorn_ex_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_int,
                                   POINTER(SinkhornOutputs), c_void_p, c_size_t, c_void_p]),
    "om_match_ex_workspace_bytes": (c_size_t, [POINTER(MatchParams)]),
    "om_match_pairs_ex": (c_int, [POINTER(MatchParams), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, POINTER(SinkhornOutputs), c_void_p, c_size_t, c_void_p]),
    "om_debug_detect_stage": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p,
                                      c_void_p, c_void_p, c_size_t, c_void_p, c_int]),
    "om_debug_dense_stage": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_float,
                                     c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_int]),
    "om_debug_force_generic_stencil": (None, [c_int]),
    "om_debug_sweep_tuning": (None, [c_int, c_int]),
    "om_debug_nms_variant": (None, [c_int]),
    "om_debug_essential_variant": (None, [c_int]),
    "om_debug_match_streams": (None, [c_int]),
    "om_debug_match_binary": (None, [c_int]),
    "om_debug_dense_window": (None, [c_int]),
    "om_debug_band_rows": (None, [c_int]),
    "om_debug_score_variant": (None, [c_int]),
    "om_debug_force_generic_sinkhorn": (None, [c_int]),
    "om_debug_sinkhorn_variant": (None, [c_int]),
    "om_debug_xl_reverse": (None, [c_int]),
    "om_debug_sinkhorn_trace": (None, [c_void_p]),
    "om_debug_hy_max_clusters": (c_int, [c_int]),
}


def lib() -> ctypes.CDLL:
    """Load (building in-tree first if needed) libom_b200.so."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            # build() returns at once when the binary is newer than every source and the header; a stale binary (edited
            # kernels, changed struct layout) is rebuilt instead of being loaded silently.  Without nvcc an existing
            # binary is used as it is (the GPU box gets the binary built in the authoring container).
            try:
                path = _build.build()      # raises if nvcc is missing: no CPU fallback exists
            except RuntimeError:
                path = _build.LIB_PATH
                if not os.path.exists(path):
                    raise
            handle = ctypes.CDLL(path)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)   # AttributeError if the library lacks a declared symbol
                fn.restype = res
                fn.argtypes = args
            _lib = handle
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().om_error_string(status).decode()
        raise RuntimeError(f"{what} failed: {msg} (status {status})")


def use_device(index: int) -> None:
    """Kept for callers that drive the C ABI by hand: every compute entry point already runs on the device that owns its
    first pointer argument and restores the caller's current device (DeviceScope in csrc/common.cuh)."""
    check(lib().om_set_device(int(index)), "om_set_device")


def launch_count() -> int:
    return int(lib().om_launch_count())
