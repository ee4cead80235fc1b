#!/usr/bin/env python
"""Turn ncu CSV exports into the short summaries committed under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/x_launches_raw.csv > profiles/x_launches.md
    python tools/summarize_ncu.py full     gpurun_out/x_full_raw.csv     > profiles/x_full.md

`launches`: the `--metrics gpu__time_duration.sum` pass of a bench command: per kernel name the
number of launches, total / mean duration and the share of all GPU time of the command.
`full`: the `--page raw --csv` export of an `ncu --set full` capture: the metrics the roofline
discussion needs, one column per captured launch.
"""
import csv
import re
import sys
from collections import OrderedDict

KEY = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
]


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name if len(name) < 90 else name[:87] + "..."


def launches(path: str) -> None:
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    total = 0.0
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
        k = short(r[ik])
        c = agg.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += v
        total += v
    print(f"source: {path}\n\ntotal GPU time of the command: {total/1e3:.3f} ms over {sum(c[0] for c in agg.values())} launches\n")
    print("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t:.1f} | {t/n:.2f} | {100*t/total:.1f}% |")


def full(path: str) -> None:
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik = hdr.index("Kernel Name")
    print(f"source: {path} (ncu --set full --clock-control none, one column per captured launch)\n")
    print("| metric | unit | " + " | ".join(f"`{short(r[ik])}` grid {r[hdr.index('Grid Size')]}" for r in data) + " |")
    print("|---|---|" + "---|" * len(data))
    for m in KEY:
        if m in hdr:
            j = hdr.index(m)
            print(f"| {m} | {units[j]} | " + " | ".join(r[j] for r in data) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
