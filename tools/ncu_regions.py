#!/usr/bin/env python3
"""Split a kernel's ncu SASS page into barrier-delimited regions and print sample share + opcode mix.
usage: tools/ncu_regions.py report.ncu-rep kernel_regex"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several launches may be present: take the first kernel block
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
b = blocks[0]
hdr, data = b["rows"][0], b["rows"][1:]
si, so, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
tot = sum(int(r[si] or 0) for r in data)
print(b["name"][:90], "samples", tot, "sass instr", len(data))
def op(r):
    t = r[so].split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
start, acc, regs = 0, 0, []
for i, r in enumerate(data):
    acc += int(r[si] or 0)
    o = op(r)
    if o in ("BAR", "UCGABAR_ARV", "UCGABAR_WAIT"):
        regs.append((start, i, acc)); start, acc = i + 1, 0
regs.append((start, len(data) - 1, acc))
for a, e, s in regs:
    ops = {}
    for r in data[a:e + 1]:
        ops[op(r)] = ops.get(op(r), 0) + int(r[ie] or 0)
    top = sorted(ops.items(), key=lambda x: -x[1])[:8]
    print(f"[{a:4d}-{e:4d}] {100 * s / max(tot, 1):5.1f}%  warp-instr {sum(ops.values()) // 1000:7d}k :", " ".join(f"{k}:{v // 1000}k" for k, v in top))
