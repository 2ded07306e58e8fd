#!/usr/bin/env python3
"""Per-kernel SASS opcode histogram of libom_b200.so (cuobjdump -sass): the Blackwell evidence (UTCHMMA = tcgen05.mma,
LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor loads, UBLKCP = bulk copies, SYNCS = mbarrier, FFMA2/FADD2/FMUL2 = packed f32x2).
usage: python tools/sass_histogram.py [lib.so] > profiles/r2_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "onnx_image_processing_b200", "_lib", "libom_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ("UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTCBAR", "UTCATOM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UCGABAR", "FFMA2", "FADD2", "FMUL2",
       "FFMA", "FADD", "FMUL", "MUFU", "SHFL", "LDS", "STS", "LDG", "STG", "LDGSTS", "ATOMS", "ATOMG", "RED", "BAR", "HMMA", "REDUX", "VIMNMX", "FMNMX")
kern, hist = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "").replace("void ", "").replace("om::", "")
        kern = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
        if m.group(1) in ("UTMALDG", "UBLKCP", "SYNCS", "MUFU", "LDS", "STS", "LDG", "STG"):
            hist[kern][m.group(1) + m.group(2)] += 0   # keep the plain count; variants listed below
print(f"# SASS opcode histogram per kernel, {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a); static instruction counts")
print("# tcgen05.mma -> UTCHMMA, tcgen05.ld/st -> LDTM/STTM, tcgen05.commit -> UTCBAR, cp.async.bulk.tensor -> UTMALDG, cp.async.bulk -> UBLKCP,")
print("# mbarrier -> SYNCS, packed f32x2 arithmetic -> FFMA2 / FADD2 / FMUL2")
tot = collections.Counter()
for k, c in hist.items():
    n = sum(c.values())
    keys = [f"{o}={c[o]}" for o in KEY if c.get(o)]
    print(f"{k[:110]}\n    total={n}  " + " ".join(keys))
    tot.update(c)
print("\n# whole library: " + " ".join(f"{o}={tot[o]}" for o in KEY if tot.get(o)))
