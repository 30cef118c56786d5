#!/usr/bin/env python3
"""Turn the raw ncu outputs of a gpurun call into the small text summaries kept under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_big.csv profiles/r01_launches_big.md
    python tools/summarize_ncu.py full gpurun_out/prof_stream.ncu-rep profiles/r01_gemv_stream_full.md
"""
import collections
import csv
import re
import subprocess
import sys


def launches(src, dst):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
        key = (re.sub(r"\(.*", "", row["Kernel Name"])[:70], row.get("Grid Size", ""), row.get("Block Size", ""))
        agg[key][0] += 1
        agg[key][1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none): {n} launches, {tot:.1f} us total\n\n")
        f.write("Per-launch times are cold-cache and serialised by ncu: read the SHARES, not the absolutes.\n\n")
        f.write("| share | total us | launches | avg us | kernel | grid | block |\n|---|---|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {100 * v[1] / tot:5.1f}% | {v[1]:9.1f} | {v[0]} | {v[1] / v[0]:7.2f} | `{k[0]}` | {k[1]} | {k[2]} |\n")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_not_selected",
        "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_membar", "smsp__inst_executed_op_local_ld.sum",
        "smsp__inst_executed_op_local_st.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none: {len(data)} captured launches from {src}\n\n")
        for d in data:
            f.write(f"## `{d[idx['Kernel Name']][:110]}`\n\n| metric | value | unit |\n|---|---|---|\n")
            for w in WANT:
                if w in idx:
                    f.write(f"| {w} | {d[idx[w]]} | {units[idx[w]]} |\n")
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
