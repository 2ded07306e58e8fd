#!/usr/bin/env python3
"""Time of the essential-matrix head (om_essential_matrix_f32) for a batch of Sinkhorn matrices."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_image_processing_b200 import _ops, _native
for clustered in (0, 1):
  _native.lib().om_debug_essential_variant(clustered)
  print("8-CTA cluster per pair" if clustered else "one CTA per pair")
  for B, K in ((1, 512), (64, 512), (64, 128), (8, 2048)):
      g = torch.Generator().manual_seed(B + K)
      P = torch.rand(B, K + 1, K + 1, generator=g).cuda()
      pts = torch.rand(B, K, 2, generator=g).cuda()
      valid = torch.ones(B, K, dtype=torch.bool, device="cuda")
      for _ in range(3): _ops.essential_matrix(P, pts, pts, valid, valid, 3, 30, 10)
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      torch.cuda.synchronize(); a.record()
      for _ in range(20): _ops.essential_matrix(P, pts, pts, valid, valid, 3, 30, 10)
      b.record(); torch.cuda.synchronize()
      print(f"B={B} K={K}: {a.elapsed_time(b) / 20 * 1000:.1f} us per call, {a.elapsed_time(b) / 20 / B * 1000:.2f} us per pair")
