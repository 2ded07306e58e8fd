"""Random-shape sweep of the streaming Sinkhorn kernels (forced at every size, variant 9) against the float64 oracle."""
import sys, torch, random
sys.path.insert(0, ".")
import onnx_image_processing_b200 as om
from onnx_image_processing_b200 import _native
from oracle import oracle as O
from tests import parity as PR
lib = _native.lib()
random.seed(3)
cases = [(5, 2147, 32, 0.3), (2147, 3, 64, 0.2), (1500, 1500, 96, 0.1), (1, 1025, 32, 1.0), (1025, 1, 160, 0.5), (2047, 2049, 64, 0.07),
         (33, 17, 32, 0.05), (1100, 1300, 512, 0.05)]
for _ in range(6):
    cases.append((random.randint(1, 2147), random.randint(1, 2147), 32 * random.randint(1, 8), random.choice([0.05, 0.1, 0.5, 1.0])))
bad = 0
for (N, M, D, eps) in cases:
    B = random.randint(1, 3)
    g = torch.Generator().manual_seed(N * 31 + M)
    d1 = torch.nn.functional.normalize(torch.randn(B, N, D, generator=g), dim=-1)
    pick = (torch.randperm(max(N, M), generator=g) % N)[:M]
    d2 = torch.nn.functional.normalize(d1[:, pick] + 0.2 * torch.randn(B, M, D, generator=g), dim=-1)
    ref = O.sinkhorn(d1.double(), d2.double(), 20, eps, 1.0).float()
    lib.om_debug_sinkhorn_variant(9)
    got = om.SinkhornMatcher(20, eps, 1.0).cuda()(d1.cuda(), d2.cuda()).cpu()
    lib.om_debug_sinkhorn_variant(0)
    m = PR.prob_metrics(got, ref)
    ok = PR.probs_ok(m)
    bad += (not ok)
    print(N, M, D, eps, B, "ok" if ok else "FAIL", {k: (round(v, 8) if isinstance(v, float) else v) for k, v in m.items()})
print("failures", bad)
