#!/usr/bin/env python3
"""One line per profiled launch of an .ncu-rep: duration, DRAM bytes, occupancy, issue rate, registers.
usage: tools/ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv, io, subprocess, sys

WANT = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("lts__t_bytes.sum", "l2_bytes"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed.sum", "warp_inst"), ("launch__registers_per_thread", "regs"),
        ("launch__occupancy_limit_shared_mem", "lim_smem"), ("launch__occupancy_limit_registers", "lim_regs"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block")]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    print("==", rep)
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        parts = []
        for key, short in WANT:
            if key in h:
                i = h.index(key)
                parts.append(f"{short}={r[i]}{units[i] if short in ('us','dram_rd','dram_wr','l2_bytes') else ''}")
        print(name[:60].ljust(60), " ".join(parts))
