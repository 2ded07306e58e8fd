"""BAD descriptors: drop-ins for pytorch_model/descriptor/bad.py (BADDescriptor :14-218,
extract_descriptors_at_keypoints[_subpixel] :221-333, SparseBAD :336-576)."""
import torch
from torch import nn

from .. import _ops
from .bad_params import _get_bad_learned_params


def _register_tables(mod: nn.Module, num_pairs: int) -> None:
    box, thr = _get_bad_learned_params(num_pairs)
    mod.register_buffer("offset_x1", box[:, 0] - 16.0)      # bad.py:33-38 / :405-410
    mod.register_buffer("offset_x2", box[:, 1] - 16.0)
    mod.register_buffer("offset_y1", box[:, 2] - 16.0)
    mod.register_buffer("offset_y2", box[:, 3] - 16.0)
    mod.register_buffer("radii", box[:, 4].to(torch.int64))
    mod.register_buffer("thresholds", thr)


def _register_bank(mod: nn.Module, num_pairs: int) -> None:
    """radius_select / box_kernel_bank: unused by the CUDA kernels, kept so state_dicts interchange."""
    R = int(mod.radii.max())
    mod.max_radius = R
    sel = torch.zeros(R + 1, num_pairs)
    sel[mod.radii, torch.arange(num_pairs)] = 1.0
    mod.register_buffer("radius_select", sel)
    c = torch.arange(-R, R + 1, dtype=torch.float32)
    gy, gx = torch.meshgrid(c, c, indexing="ij")
    rv = torch.arange(R + 1, dtype=torch.float32).view(-1, 1, 1)
    masks = ((gy.abs() <= rv) & (gx.abs() <= rv)).to(torch.float32)
    mod.register_buffer("box_kernel_bank", (masks / ((2.0 * rv + 1.0) ** 2).clamp_min(1.0)).unsqueeze(1))


class _PairTable:
    """``_pair_table``: (P,6) rows {ox1, ox2, oy1, oy2, radius, threshold}, what the C ABI consumes.  It is DERIVED from
    the module's buffers -- the ones the reference's forward reads (bad.py:33-38 for the dense module, the ``*_v`` views
    and ``radius_select`` of bad.py:413-423 for SparseBAD) -- and rebuilt whenever one of them changed (load_state_dict,
    in-place edits, .to(device)), so a loaded state_dict is what the kernels use.  Not a buffer: state_dict keys stay
    exactly the reference's."""

    def _pair_sources(self):
        return (self.offset_x1, self.offset_x2, self.offset_y1, self.offset_y2, self.radii, self.thresholds)

    def _pair_columns(self):
        return [self.offset_x1, self.offset_x2, self.offset_y1, self.offset_y2, self.radii.float(), self.thresholds]

    @property
    def _pair_table(self) -> torch.Tensor:
        key = tuple((b.data_ptr(), b._version, str(b.device)) for b in self._pair_sources())
        cache = self.__dict__.get("_pair_cache")
        if cache is None or cache[0] != key:
            with torch.no_grad():
                t = torch.stack([c.reshape(-1).float() for c in self._pair_columns()], dim=1).contiguous()
            cache = (key, t)
            self.__dict__["_pair_cache"] = cache
        return cache[1]


class BADDescriptor(_PairTable, nn.Module):
    """Dense BAD map (B,1,H,W) -> (B,num_pairs,H,W) (bad.py:14-218, non-oriented path)."""

    def __init__(self, num_pairs: int = 256, binarize: bool = False, soft_binarize: bool = True,
                 temperature: float = 10.0) -> None:
        super().__init__()
        self.num_pairs = num_pairs
        self.binarize = binarize
        self.soft_binarize = soft_binarize
        self.temperature = temperature
        _register_tables(self, num_pairs)
        self.register_buffer("area", ((2.0 * self.radii.float() + 1.0) ** 2).view(-1, 1, 1))
        _register_bank(self, num_pairs)

    def forward(self, x: torch.Tensor, orientation: torch.Tensor | None = None) -> torch.Tensor:
        if orientation is not None:
            # bad.py:112-187 (per-pixel rotated dense BAD) is only reached from the AKAZE models,
            # which are outside this build's hot path.
            raise NotImplementedError("dense oriented BAD is not part of the B200 hot path; use SparseBAD")
        return _ops.dense_bad(x, self._pair_table, _ops.desc_mode(self.binarize, self.soft_binarize),
                              float(self.temperature))


def extract_descriptors_at_keypoints(descriptor_map: torch.Tensor, keypoints: torch.Tensor) -> torch.Tensor:
    """Integer gather (B,D,H,W),(B,N,2) -> (B,N,D) (bad.py:221-274)."""
    return _ops.gather_descriptors(descriptor_map, keypoints, False)


def extract_descriptors_at_keypoints_subpixel(descriptor_map: torch.Tensor, keypoints: torch.Tensor) -> torch.Tensor:
    """Bilinear gather with border clamping (bad.py:277-333)."""
    return _ops.gather_descriptors(descriptor_map, keypoints, True)


class SparseBAD(_PairTable, nn.Module):
    """BAD descriptors at keypoints only, optionally rotated by a per-pixel orientation map
    (bad.py:336-576)."""

    def __init__(self, num_pairs: int = 256, binarize: bool = False, soft_binarize: bool = True,
                 temperature: float = 10.0, normalize_descriptors: bool = True, sampling_mode: str = "nearest"):
        super().__init__()
        if num_pairs not in (256, 512):
            raise ValueError(f"num_pairs must be 256 or 512 to use learned BAD patterns, got {num_pairs}")
        if sampling_mode not in ("nearest", "bilinear"):
            raise ValueError(f"sampling_mode must be 'nearest' or 'bilinear', got {sampling_mode}")
        self.num_pairs = num_pairs
        self.binarize = binarize
        self.soft_binarize = soft_binarize
        self.temperature = temperature
        self.normalize_descriptors = normalize_descriptors
        self.sampling_mode = sampling_mode
        _register_tables(self, num_pairs)
        for name in ("offset_y1", "offset_x1", "offset_y2", "offset_x2", "thresholds"):   # bad.py:413-417
            self.register_buffer(name + "_v", getattr(self, name).view(1, 1, -1))
        _register_bank(self, num_pairs)

    def _pair_sources(self):                 # what SparseBAD.forward of the reference reads (bad.py:520-525, :554-559)
        return (self.offset_x1_v, self.offset_x2_v, self.offset_y1_v, self.offset_y2_v, self.radius_select, self.thresholds_v)

    def _pair_columns(self):
        return [self.offset_x1_v, self.offset_x2_v, self.offset_y1_v, self.offset_y2_v,
                self.radius_select.argmax(dim=0).float(), self.thresholds_v]

    def _mode(self) -> int:
        return _ops.desc_mode(self.binarize, self.soft_binarize)

    def forward(self, image: torch.Tensor, keypoints: torch.Tensor,
                orientation: torch.Tensor | None = None) -> torch.Tensor:
        theta_mode = _ops.THETA_NONE if orientation is None else _ops.THETA_MAP
        return _ops.sparse_bad(image, keypoints, self._pair_table, self._mode(), float(self.temperature),
                               bool(self.normalize_descriptors), _ops.sampling_code(self.sampling_mode), theta_mode,
                               orientation, None)
