#!/usr/bin/env python
"""Print the metrics we track from an .ncu-rep (run here, no GPU needed):  python profiles/ncu_summary.py rep [filter]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = sys.argv[2:] or ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_subpipe",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "smsp__average_warp", "smsp__average_warps_issue_stalled", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "tensor"]
for r in rows[2:]:
    for i, h in enumerate(hdr):
        if any(h == w or (len(w) > 12 and h.startswith(w)) or (w == "tensor" and "tensor" in h) for w in want):
            print(f"{h:95s} {r[i]:>22s} {units[i]}")
    print("---")
