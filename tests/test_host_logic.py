"""CPU-only tests: the C-ABI library loads and exports everything include/om_b200.h declares, the
nn.Module mirror has the reference's constructor contract, sharding/gather host logic works over
gloo with world_size 2.  No compute entry point is called (there is no GPU here)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

import onnx_image_processing_b200 as om
from onnx_image_processing_b200 import _native
from onnx_image_processing_b200.host_pipeline import shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "om_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(om_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _native.lib()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in om_b200.h but not exported"
        assert n in _native.SIGNATURES, f"{n} has no ctypes signature"
    assert lib.om_version() >= 100
    assert lib.om_error_string(0) == b"ok"
    assert b"workspace" in lib.om_error_string(4)


def test_argument_errors_without_gpu():
    """Validation happens before any CUDA call, so these run on a CPU-only box."""
    lib = _native.lib()
    null = ctypes.c_void_p(0)
    assert lib.om_shi_tomasi_score_f32(null, 1, 8, 8, 3, null, null) == 1           # OM_ERR_NULL
    assert lib.om_topk_workspace_bytes(2, 480, 640, 512) >= 2 * 480 * 640 * 8
    assert lib.om_sinkhorn_workspace_bytes(1, 512, 512, 256) > 0
    assert lib.om_dense_bad_workspace_bytes(1, 480, 640) > 8 * 480 * 640 * 4
    prm = _native.MatchParams(0, 4, 480, 640, 512, 3, 3, 7, 0.0, 256, 0, 10.0, 1, 0, 15, 20, 1.0, 1.0, 0)
    assert lib.om_match_workspace_bytes(ctypes.byref(prm)) > 0
    prm.P = 300
    assert lib.om_match_workspace_bytes(ctypes.byref(prm)) == 0                      # num_pairs must be 256/512


def test_bad_table_matches_packaged_data():
    lib = _native.lib()
    for n in (256, 512):
        boxes = (ctypes.c_byte * (n * 5))()
        thr = (ctypes.c_float * n)()
        assert lib.om_bad_table(n, boxes, thr) == 0
        sb = om.SparseBAD(num_pairs=n)
        got = torch.tensor(list(boxes), dtype=torch.float32).view(n, 5)
        assert torch.equal(got[:, 0], sb.offset_x1) and torch.equal(got[:, 1], sb.offset_x2)
        assert torch.equal(got[:, 2], sb.offset_y1) and torch.equal(got[:, 3], sb.offset_y2)
        assert torch.equal(got[:, 4].long(), sb.radii)
        assert torch.equal(torch.tensor(list(thr)), sb.thresholds)
        assert sb._pair_table.shape == (n, 6) and "_pair_table" not in sb.state_dict()
    assert lib.om_bad_table(128, boxes, thr) == 3


def test_constructor_errors_match_reference():
    with pytest.raises(ValueError):
        om.ShiTomasiScore(sobel_size=5)
    with pytest.raises(ValueError):
        om.ShiTomasiScore(block_size=4)
    with pytest.raises(ValueError):
        om.SparseBAD(num_pairs=128)
    with pytest.raises(ValueError):
        om.SparseBAD(sampling_mode="cubic")
    with pytest.raises(ValueError):
        om.BADDescriptor(num_pairs=100)
    with pytest.raises(ValueError):
        om.SinkhornMatcher(iterations=0)
    with pytest.raises(ValueError):
        om.SinkhornMatcher(epsilon=0.0)
    with pytest.raises(ValueError):
        om.SinkhornMatcher(distance_type="cosine")
    with pytest.raises(ValueError):
        om.AngleEstimator(patch_size=14)
    with pytest.raises(ValueError):
        om.AngleEstimator(sigma=0.0)


def test_module_defaults_and_state_dict_names():
    m = om.ShiTomasiSparseBADSinkhornMatcher(512)
    assert m.border_margin == 7 and m.nms_radius == 3 and m.matcher.iterations == 20
    assert om.ShiTomasiSparseBADSinkhornMatcher(8, border_margin=0).border_margin == 0
    keys = set(m.state_dict())
    assert {"corner_detector.sobel_xy", "corner_detector.sum_kernel_grouped", "descriptor.offset_x1",
            "descriptor.radii", "descriptor.thresholds_v", "descriptor.radius_select",
            "descriptor.box_kernel_bank"} <= keys
    a = om.ShiTomasiAngleSparseBADSinkhornMatcher(64)
    assert a.detector.shi_tomasi.block_size == 5
    assert "detector.angle_estimator.moment_kernels" in a.state_dict()
    d = om.ShiTomasiBADSinkhornMatcher(64)
    assert "detector.descriptor.area" in d.state_dict()
    assert tuple(m.corner_detector.sobel_xy[1, 0, 0].tolist()) == (-1.0, -2.0, -1.0)


def test_no_cpu_fallback():
    m = om.ShiTomasiSparseBADSinkhornMatcher(16)
    with pytest.raises((NotImplementedError, RuntimeError)):
        m(torch.zeros(1, 1, 32, 32), torch.zeros(1, 1, 32, 32))
    with pytest.raises((NotImplementedError, RuntimeError)):
        om.SinkhornMatcher()(torch.zeros(1, 4, 8), torch.zeros(1, 4, 8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "onnx_image_processing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} mentions the oracle"


def test_shard_bounds_cover_batch():
    for n in (0, 1, 7, 64, 1000):
        for w in (1, 2, 3, 8):
            parts = [shard_bounds(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from onnx_image_processing_b200.host_pipeline import shard_bounds, gather_to_rank0
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
r = dist.get_rank()
total = 7
lo, hi = shard_bounds(total, r, 2)
# stand-in for per-shard matcher outputs: values identify the global pair index
k = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1, 1).expand(-1, 4, 2).contiguous()
p = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1, 1).expand(-1, 5, 5).contiguous()
out = gather_to_rank0([k, p])
if r == 0:
    assert out[0].shape == (total, 4, 2) and out[1].shape == (total, 5, 5)
    assert torch.equal(out[0][:, 0, 0], torch.arange(total, dtype=torch.float32))
    assert torch.equal(out[1][:, 0, 0], torch.arange(total, dtype=torch.float32))
    print("GATHER_OK")
else:
    assert out is None
dist.destroy_process_group()
"""


def test_two_rank_shard_and_gather_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK" in outs[0]


def test_bench_reference_arm_prints_contract_line():
    """bench.py --impl reference runs the oracle port on the CPU and prints the JSON contract."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "sparse",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_numa_binding_helper_is_harmless_without_topology():
    """bind_host_to_gpu: parses sysfs CPU lists; without a GPU / NUMA topology it reports and changes nothing."""
    import os
    from onnx_image_processing_b200.host_pipeline import _cpu_list, bind_host_to_gpu
    assert _cpu_list("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert _cpu_list("") == set()
    before = os.sched_getaffinity(0)
    info = bind_host_to_gpu(0)
    assert info["bound"] is False or info.get("cpus", 0) > 0
    if not info["bound"]:
        assert os.sched_getaffinity(0) == before
