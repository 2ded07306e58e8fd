"""Export-default configuration (K=1024, 512 pairs, hard binarisation, eps 0.05, NMS radius 5), batch 64: step and stage times."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda", 0)
rows = bench.measure_other_configs(dev, 8)
for r in rows:
    print(r["config"][:60], r["pairs_per_step"], f'{r["ms_per_step"]:.3f} ms', round(r["pairs_per_s"]), r.get("stage_ms_timed_alone"))
