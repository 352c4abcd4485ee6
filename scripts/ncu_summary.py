#!/usr/bin/env python
"""One line per captured kernel from an .ncu-rep: python scripts/ncu_summary.py file.ncu-rep [out.md]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"), ("lts__t_bytes.sum", "l2_bytes"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lsu_smem_wavefronts"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%")]
idx = [(hdr.index(n), s) for n, s in want if n in hdr]
units = rows[1]
lines = ["| " + " | ".join(s + (f" [{units[i]}]" if units[i] else "") for i, s in idx) + " |", "|" + "---|" * len(idx)]
for r in rows[2:]:
    vals = []
    for i, s in idx:
        v = r[i]
        if s == "kernel":
            v = v.split("(")[0].replace("void pv::<unnamed>::", "").replace("pv::sl::", "")[:40]
        vals.append(v)
    lines.append("| " + " | ".join(vals) + " |")
text = "\n".join(lines) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(f"# ncu --set full summary of {rep}\n\n" + text)
print(text)
