"""Compact per-kernel summary of an `ncu --page raw --csv` export: duration, DRAM bytes, throughput percentages.
    python scripts/summarize_ncu_raw.py gpurun_out/prof_tile_raw.csv > profiles/rNN_ncu_full_tile_kernels.csv"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "smsp__inst_executed.sum"]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keys = [k for k in KEYS if k in hdr]
w = csv.writer(sys.stdout)
w.writerow(["kernel"] + [f"{k} [{units[hdr.index(k)]}]" for k in keys])
for r in rows[2:]:
    w.writerow([r[hdr.index("Kernel Name")][:70]] + [r[hdr.index(k)] for k in keys])
