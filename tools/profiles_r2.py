#!/usr/bin/env python3
"""profiles/r2_* from what tools/gpu_profile_r2.sh left in gpurun_out/ (run here after the GPU call):
launch lists, one-line-per-launch summaries of the `--set full` captures, DRAM bytes per launch per kernel
(bench.py reads r2_dram_traffic.json for roofline.traffic) and `details` excerpts of the largest kernels."""
import csv, json, os, re, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAGS = ("r2_dense_b64", "r2_sparse_b64")
traffic = {}
for tag in TAGS:
    for src, dst in ((f"launches_{tag}.csv", f"r2_launches_{tag[3:]}.csv"), (f"ncu_full_summary_{tag}.txt", f"r2_ncu_full_summary_{tag[3:]}.txt")):
        if os.path.exists(os.path.join(G, src)):
            txt = open(os.path.join(G, src), errors="ignore").read().replace("gpurun_out/", "")
            open(os.path.join(P, dst), "w").write(txt)
    raw = os.path.join(G, f"ncu_raw_{tag}.csv")
    if not os.path.exists(raw):
        continue
    rows = list(csv.reader(open(raw, errors="ignore")))
    h, units = rows[0], rows[1]
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}
    def col(r, name):
        i = h.index(name)
        return float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        short = re.sub(r"<.*", "", name.replace("<unnamed>::", "").split("(")[0]).split("::")[-1].split()[-1].strip()
        t = traffic.setdefault(short, {"dram_bytes_per_launch": 0.0, "ncu_us_per_launch": 0.0, "launches": 0})
        t["dram_bytes_per_launch"] += col(r, "dram__bytes_read.sum") + col(r, "dram__bytes_write.sum")
        t["ncu_us_per_launch"] += col(r, "gpu__time_duration.sum")
        t["launches"] += 1
for t in traffic.values():
    t["dram_bytes_per_launch"] /= t["launches"]
    t["ncu_us_per_launch"] /= t["launches"]
traffic["_how"] = ("ncu --set full --clock-control none, python tools/profile_step.py {dense,sparse} 64 2 (batch 64 pairs, 480x640, k=512); "
                   "dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the launches captured (cold L2: every launch is replayed alone)")
json.dump(traffic, open(os.path.join(P, "r2_dram_traffic.json"), "w"), indent=1)

KEEP = ("Duration", "Memory Throughput", "DRAM Throughput", "L2 Cache Throughput", "L1/TEX Cache Throughput", "Executed Ipc Active", "Issue Slots Busy",
        "L1/TEX Hit Rate", "L2 Hit Rate", "No Eligible", "Warp Cycles Per Issued Instruction", "Registers Per Thread",
        "Dynamic Shared Memory Per Block", "Static Shared Memory Per Block", "Block Limit", "Theoretical Occupancy",
        "Achieved Occupancy", "Cluster Size", "Max Active Clusters", "Compute (SM) Throughput", "Grid Size", "Block Size", "Waves Per SM",
        "Eligible Warps Per Scheduler", "Mem Pipes Busy", "Executed Instructions")
for tag in TAGS:
    det = os.path.join(G, f"ncu_details_{tag}.txt")
    if not os.path.exists(det):
        continue
    blocks, cur = {}, None
    for l in open(det, errors="ignore"):
        s = l.strip()
        m = re.search(r"([a-z0-9_]+_kernel)[<(]", s)
        if m and "Context" in s and "Device" in s:
            cur = m.group(1)
            if cur in blocks:
                cur = None           # first launch of each kernel only
            else:
                blocks[cur] = ["  " + s[:170]]
            continue
        if cur and any(s.startswith(k) for k in KEEP) and ("    " + s) not in blocks[cur]:
            blocks[cur].append("    " + s)
    for k, lines in blocks.items():
        if k in ("prefix_cols_kernel", "prefix_rows_kernel", "sparse_bad_kernel"):
            continue
        fn = os.path.join(P, f"r2_ncu_details_{k.replace('_kernel', '')}.txt")
        if tag.endswith("sparse_b64") and os.path.exists(fn):
            continue
        open(fn, "w").write(f"# ncu --set full, details page, first captured launch of {k} ({tag[3:]} step, batch 64)\n" + "\n".join(lines) + "\n")
for src, dst in (("hy_k512.txt", "r2_sinkhorn_hy_trace_k512_b64.txt"), ("hy_k1024.txt", "r2_sinkhorn_hy_trace_k1024_b64.txt"),
                 ("hy_probe.txt", "r2_sinkhorn_variants.txt")):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
print(json.dumps({k: v for k, v in traffic.items() if k != "_how"}, indent=1))
