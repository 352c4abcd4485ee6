#!/usr/bin/env python
"""Pipe / issue / stall summary per captured kernel from an `ncu --set full` report:
python scripts/ncu_detail.py file.ncu-rep [out.md] [title]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
M = [("gpu__time_duration.sum", "duration"),
     ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
     ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
     ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
     ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
     ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
     ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX (incl. shared memory) %"),
     ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
     ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
     ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
     ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
     ("smsp__inst_executed.sum", "warp instructions executed"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
     ("launch__registers_per_thread", "registers / thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
     ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected (eligible, lost arbitration)"),
     ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
     ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard (shared memory / SFU)"),
     ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (global / TMEM loads)"),
     ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
     ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall: dispatch"),
     ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: MIO throttle"),
     ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
     ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall: sleeping (mbarrier try_wait)")]
idx = [(hdr.index(n), s) for n, s in M if n in hdr]
ik = hdr.index("Kernel Name")
out = [f"# {sys.argv[3] if len(sys.argv) > 3 else rep}", "", f"`ncu --set full --clock-control none` ({rep.split('/')[-1]}); one column per captured launch.", ""]
kern = [r for r in rows[2:] if len(r) > ik]
out.append("| metric | " + " | ".join(r[ik].split("(")[0].replace("void pv::<unnamed>::", "").replace("pv::sl::", "")[:36] for r in kern) + " |")
out.append("|---|" + "---|" * len(kern))
for i, s in idx:
    out.append(f"| {s} [{units[i]}] | " + " | ".join(r[i] for r in kern) + " |")
text = "\n".join(out) + "\n"
if len(sys.argv) > 2 and sys.argv[2] != "-":
    open(sys.argv[2], "w").write(text)
print(text)
