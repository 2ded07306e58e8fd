#!/usr/bin/env python3
"""profiles/ files from the ncu reports in gpurun_out/ (run here, after the GPU call that captured them):
r1_ncu_full_summary.txt, r1_dram_traffic.json and the "details" excerpts of the largest kernels."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPS = [os.path.join(ROOT, "gpurun_out", n) for n in ("prof_r1_dense_step.ncu-rep", "prof_r1_sparse_bad.ncu-rep")]
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), *REPS], capture_output=True, text=True).stdout
open(os.path.join(ROOT, "profiles", "r1_ncu_full_summary.txt"), "w").write(out.replace(ROOT + "/", ""))

traffic = {}
for rep in REPS:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    def col(r, name, scale={"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}):
        i = h.index(name)
        return float(r[i]) * scale[units[i]]
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        short = name.replace("<unnamed>::", "").replace("unnamed>::", "").split("(")[0].split("<")[0].split("::")[-1].split()[-1].strip()
        if short.startswith("reduce_kernel"):
            continue
        t = traffic.setdefault(short, {"dram_bytes_per_launch": 0.0, "ncu_us_per_launch": 0.0, "launches": 0})
        t["dram_bytes_per_launch"] += col(r, "dram__bytes_read.sum") + col(r, "dram__bytes_write.sum")
        t["ncu_us_per_launch"] += col(r, "gpu__time_duration.sum")
        t["launches"] += 1
for t in traffic.values():
    t["dram_bytes_per_launch"] /= t["launches"]
    t["ncu_us_per_launch"] /= t["launches"]
traffic["_how"] = ("ncu --set full --clock-control none, python tools/profile_step.py {dense,sparse} 64 1 (batch 64 pairs, 480x640, "
                   "k=512); dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the launches captured")
json.dump(traffic, open(os.path.join(ROOT, "profiles", "r1_dram_traffic.json"), "w"), indent=1)

KEEP = ("Duration", "Memory Throughput", "DRAM Throughput", "L2 Cache Throughput", "Executed Ipc Active", "Issue Slots Busy",
        "L1/TEX Hit Rate", "L2 Hit Rate", "No Eligible", "Warp Cycles Per Issued Instruction", "Registers Per Thread",
        "Dynamic Shared Memory Per Block", "Static Shared Memory Per Block", "Block Limit", "Theoretical Occupancy",
        "Achieved Occupancy", "Cluster Size", "Max Active Clusters", "Compute (SM) Throughput", "Grid Size", "Block Size")
for rx, fn in (("score3_sweep", "score3_sweep"), ("nms3_sweep", "nms3_sweep"), ("sinkhorn_tc", "sinkhorn_tc"),
               ("dense_at_kpts", "dense_at_kpts")):
    det = subprocess.run(["ncu", "-i", REPS[0], "--page", "details", "--kernel-name", f"regex:{rx}", "--launch-count", "1"],
                         capture_output=True, text=True).stdout.splitlines()
    lines, seen = [], set()
    for l in det:
        s = l.strip()
        if s.startswith("void") or s.startswith("om::"):
            lines.append("  " + s[:160])
        elif any(s.startswith(k) for k in KEEP) and s not in seen:
            seen.add(s)
            lines.append("    " + s)
    open(os.path.join(ROOT, "profiles", f"r1_ncu_details_{fn}.txt"), "w").write("\n".join(lines) + "\n")
print(json.dumps({k: v for k, v in traffic.items() if k != "_how"}, indent=1))
