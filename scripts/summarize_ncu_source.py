"""Condense one `ncu --set full --import-source on` capture (exported with --page raw --csv and
--page source --csv --print-source sass) into a short text file: the headline counters and the SASS
instructions that collect the most warp-stall samples. usage:
    summarize_ncu_source.py <raw.csv> <sass.csv> <out.txt> [title]"""
import csv, sys

raw_p, sass_p, out_p = sys.argv[1:4]
title = sys.argv[4] if len(sys.argv) > 4 else raw_p
raw = list(csv.reader(open(raw_p, errors="replace")))
h, u, v = raw[0], raw[1], raw[2]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
lines = [f"# {title}", "# ncu --set full --clock-control none --import-source on (one launch); counters:"]
for k in KEYS:
    if k in h:
        i = h.index(k)
        lines.append(f"  {k:84s} {v[i][:60]:>24s} {u[i]}")
rows = list(csv.reader(open(sass_p, errors="replace")))
hdr = rows[1]
i_s, i_src = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source")
data = [r for r in rows[2:] if len(r) > i_s and r[i_s].isdigit()]
tot = sum(int(r[i_s]) for r in data)
lines.append(f"# warp-stall samples: {tot} over {len(data)} SASS instructions; instructions with >= 1 % of the samples:")
for k, r in enumerate(data):
    if int(r[i_s]) * 100 >= tot:
        lines.append(f"  #{k:5d} {int(r[i_s]):6d} {100 * int(r[i_s]) / tot:5.1f}%  {r[i_src].strip()[:110]}")
open(out_p, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:12]))
