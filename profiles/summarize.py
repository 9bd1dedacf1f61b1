#!/usr/bin/env python
"""Turn ncu output (gpurun_out/, scratch) into the small tracked summaries under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_X.csv  profiles/rNN_launches_X.md
  python profiles/summarize.py full     gpurun_out/prof_X.ncu-rep   profiles/rNN_full_X.md

`launches`: the `--metrics gpu__time_duration.sum --clock-control none --csv` pass -> per-kernel
count / total / mean / min / max and each kernel's share of the captured GPU time.
`full`: one `--set full` capture -> the handful of raw metrics the roofline argument uses
(duration, DRAM bytes, L2/L1 sectors, issue utilisation, pipe utilisation, stall reasons, occupancy).
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_read.sum",
    "lts__t_sectors_op_write.sum",
    "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__grid_size",
    "launch__block_size",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_static",
    "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps",
    "smsp__inst_executed.sum",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_membar_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
    "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct",
    "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 5]
    h = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H, rows = rows[h], rows[h + 1:]
    ki, vi, ui, gi, bi = (H.index(x) for x in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
    agg = collections.OrderedDict()
    for r in rows:
        name = r[ki].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        a = agg.setdefault(name, {"n": 0, "t": 0.0, "v": [], "grid": r[gi], "block": r[bi]})
        a["n"] += 1
        a["t"] += v
        a["v"].append(v)
    total = sum(a["t"] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary of `{src}`\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: use the SHARES, "
                "not the absolute times; bench numbers come from CUDA events without a profiler).\n\n")
        f.write(f"{len(rows)} launches, {total / 1e3:.3f} ms of GPU time in total.\n\n")
        f.write("| kernel | launches | total us | share | mean us | min us | max us | grid (first) | block |\n|---|---|---|---|---|---|---|---|---|\n")
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
            f.write(f"| `{name}` | {a['n']} | {a['t']:.1f} | {100 * a['t'] / total:.1f}% | {a['t'] / a['n']:.1f} | {min(a['v']):.1f} | "
                    f"{max(a['v']):.1f} | {a['grid']} | {a['block']} |\n")
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, U, D = rows[h], rows[h + 1], rows[h + 2:]
    ki = H.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of `{src}`\n\n`ncu --set full --clock-control none --import-source on` (one kernel replayed ~40x; "
                "durations here are profiler-side, the bench value is measured separately with CUDA events).\n\n")
        f.write("| metric | unit | " + " | ".join(f"#{d[0]} `{d[ki].split('(')[0].replace('void ', '')[:40]}`" for d in D) + " |\n")
        f.write("|---|---|" + "---|" * len(D) + "\n")
        for m in FULL_METRICS:
            if m in H:
                i = H.index(m)
                f.write(f"| {m} | {U[i]} | " + " | ".join(d[i] for d in D) + " |\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
