"""Times the Sinkhorn stage beyond 1024 keypoints: streaming kernels of sinkhorn_xl.cu (variant 0) against the older generic
global-memory kernels (variant 2), with and without the backwards sweeps; compares the outputs.  CUDA events, inputs in HBM."""
import sys, torch
sys.path.insert(0, ".")
from onnx_image_processing_b200 import _native, _ops

lib = _native.lib()
dev = "cuda:0"


def run(B, K, D, eps, variants, reps=10):
    g = torch.Generator().manual_seed(1)
    d1 = torch.nn.functional.normalize(torch.randn(B, K, D, generator=g), dim=-1).to(dev)
    d2 = torch.nn.functional.normalize(d1 + 0.2 * torch.randn(B, K, D, device=dev), dim=-1)
    outs = {}
    for v, rev in variants:
        lib.om_debug_sinkhorn_variant(v)
        lib.om_debug_xl_reverse(rev)
        for _ in range(3):
            p = _ops.sinkhorn(d1, d2, 20, eps, 1.0, False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            p = _ops.sinkhorn(d1, d2, 20, eps, 1.0, False)
        e1.record()
        torch.cuda.synchronize()
        outs[(v, rev)] = p
        print(f"B={B} K={K} D={D} eps={eps} variant={v} reverse={rev}: {e0.elapsed_time(e1) / reps * 1000:.1f} us per call", flush=True)
    lib.om_debug_sinkhorn_variant(0)
    lib.om_debug_xl_reverse(1)
    vs = list(outs)
    for v in vs[1:]:
        d = (outs[vs[0]] - outs[v]).abs()
        print(f"   max |P{vs[0]} - P{v}| core {float(d[:, :K, :K].max()):.3e}  all-but-corner {float(d.flatten(1)[:, :-1].max()):.3e}"
              f"  row-sum err {float((outs[v][:, :K].sum(-1) - 1).abs().max()):.2e}")


if __name__ == "__main__":
    if "--policies" in sys.argv:
        run(8, 2048, 256, 1.0, [(0, 1), (0, 0), (0, 1 + 16), (0, 1 + 32), (0, 1 + 48), (0, 1 + 64), (0, 1 + 80), (0, 1 + 96), (0, 0 + 32), (0, 0 + 48), (0, 0 + 96)])
        sys.exit(0)
    run(8, 2048, 256, 1.0, [(2, 1), (0, 1), (0, 0), (0, 3)])
    run(8, 2048, 256, 0.05, [(2, 1), (0, 1)])
    run(2, 1100, 256, 0.1, [(2, 1), (0, 1)])
    run(3, 300, 64, 0.1, [(2, 1), (9, 1)])
    run(1, 2048, 256, 1.0, [(2, 1), (0, 1)])
