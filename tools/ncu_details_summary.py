#!/usr/bin/env python3
"""Key figures per kernel of an `ncu --page details --csv` dump."""
import collections, csv, sys
KEEP = ("Duration", "DRAM Throughput", "Memory Throughput", "L2 Cache Throughput", "Achieved Occupancy", "Theoretical Occupancy",
        "Registers Per Thread", "Executed Ipc Active", "Issue Slots Busy", "No Eligible", "Eligible Warps Per Scheduler",
        "Active Warps Per Scheduler", "Block Limit Registers", "Block Limit Shared Mem", "Block Limit Warps", "Waves Per SM",
        "L1/TEX Hit Rate", "L2 Hit Rate", "Mem Busy", "Max Bandwidth", "Warp Cycles Per Issued Instruction", "Executed Instructions",
        "Shared Memory Configuration Size", "Dynamic Shared Memory Per Block", "Compute (SM) Throughput", "Mem Pipes Busy")
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = rows[0]
ix = {n: i for i, n in enumerate(hdr)}
out = collections.OrderedDict()
for r in rows[1:]:
    if len(r) != len(hdr):
        continue
    key = (r[ix["ID"]], r[ix["Kernel Name"]].split("(")[0][-60:])
    name = r[ix["Metric Name"]]
    if name in KEEP:
        out.setdefault(key, []).append(f"{name}={r[ix['Metric Value']]}{r[ix['Metric Unit']]}")
for k, v in out.items():
    print(k)
    print("   " + "; ".join(v))
# stall reasons (rule text) if present
for r in rows[1:]:
    if len(r) == len(hdr) and r[ix.get("Rule Name", 0)] and "Stall" in r[ix.get("Rule Name", 0)]:
        print("  RULE", r[ix["ID"]], r[ix["Rule Description"]][:300] if "Rule Description" in ix else "")
