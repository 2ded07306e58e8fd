"""Times the Sinkhorn stage alone: hybrid-resident cluster kernel (variant 0) against the 8-CTA tcgen05 kernel (8, K <= 512)
or the generic global-memory kernels (2, K > 512).  CUDA events, inputs in HBM."""
import sys, torch
sys.path.insert(0, ".")
from onnx_image_processing_b200 import _native, _ops

lib = _native.lib()
dev = "cuda:0"

def run(B, K, D, eps, variants, reps=20):
    g = torch.Generator().manual_seed(1)
    d1 = torch.nn.functional.normalize(torch.randn(B, K, D, generator=g), dim=-1).to(dev)
    d2 = torch.nn.functional.normalize(d1 + 0.2 * torch.randn(B, K, D, device=dev), dim=-1)
    outs = {}
    for v in variants:
        lib.om_debug_sinkhorn_variant(v)
        for _ in range(3):
            p = _ops.sinkhorn(d1, d2, 20, eps, 1.0, False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            p = _ops.sinkhorn(d1, d2, 20, eps, 1.0, False)
        e1.record()
        torch.cuda.synchronize()
        outs[v] = p
        print(f"B={B} K={K} D={D} eps={eps} variant={v}: {e0.elapsed_time(e1) / reps * 1000:.1f} us per call", flush=True)
    lib.om_debug_sinkhorn_variant(0)
    vs = list(outs)
    for v in vs[1:]:
        d = (outs[vs[0]] - outs[v]).abs()
        print(f"   max |P{vs[0]} - P{v}| core {float(d[:, :K, :K].max()):.3e}  all-but-corner {float(d.flatten(1)[:, :-1].max()):.3e}")

if __name__ == "__main__":
    run(64, 512, 256, 1.0, [0, 8])
    run(64, 512, 256, 0.05, [0, 8])
    run(1, 512, 256, 1.0, [0, 8])
    run(64, 1024, 512, 0.05, [0, 2], reps=5)
    run(8, 1024, 512, 0.05, [0, 2], reps=5)
