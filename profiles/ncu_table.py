"""Markdown table of selected metrics from `ncu -i report.ncu-rep --page raw --csv`.  usage: ncu_table.py raw.csv"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
print("| metric (unit) | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |")
print("|---|" + "---|" * len(data))
print("| kernel | " + " | ".join(r[ki].split("(")[0].replace("void <unnamed>::", "")[:40] for r in data) + " |")
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        print(f"| {w} ({units[i]}) | " + " | ".join(r[i] for r in data) + " |")
