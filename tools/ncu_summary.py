#!/usr/bin/env python
"""ncu_summary.py -- condense an Nsight Compute report into the numbers DESIGN.md / bench.py cite.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_c2.md [--traffic-key c2]

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU), writes a markdown
table per profiled kernel and, with --traffic-key, updates profiles/traffic.json with the
per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) that bench.py reports as
`roofline.traffic`.
"""
from __future__ import annotations

import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), CTAs/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), CTAs/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("lts__t_sectors_op_red.sum", "L2 red sectors"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 throughput % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 LSU wavefronts % of peak"),
    ("l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "L1 global-load bytes"),
    ("l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum", "L1 red accesses"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global-load requests (warp-wide instructions)"),
    ("l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "global-load wavefronts"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "global-red requests"),
    ("l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum", "global-red wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts.sum", "LSU data-pipe wavefronts (all)"),
    ("smsp__inst_executed_op_global_red.sum", "global-red warp instructions"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instruction"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe ALU %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe FMA %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "pipe FP64 %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe XU %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe LSU %"),
]


def load(report: str):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], check=True, capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def num(s: str) -> float:
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
              "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out_md")
    ap.add_argument("--traffic-key", default=None)
    ap.add_argument("--title", default=None)
    ap.add_argument("--command", default=None, help="the command the capture ran (recorded verbatim)")
    args = ap.parse_args()
    hdr, units, rows = load(args.report)
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {args.title or os.path.basename(args.report)}", ""]
    if args.command:
        lines += [f"Capture command: `{args.command}`", ""]
    lines += ["Source: `ncu --set full --clock-control none --import-source on`; figures are per launch, under the",
              "profiler (cold caches, serialised) -- bench.py times the same kernels live with CUDA events.", ""]
    traffic = {}
    l1pct = {}
    wpr = {}
    for r in rows:
        name = r[col["Kernel Name"]]
        short = re.sub(r"^void\s+", "", name)
        short = re.sub(r"\(.*$", "", short).replace("unnamed>::", "").replace("<unnamed>::", "")
        lines += [f"## `{short}`", "", "| metric | value |", "|---|---|"]
        for key, label in METRICS:
            if key in col:
                lines.append(f"| {label} (`{key}`) | {r[col[key]]} {units[col[key]]} |")
        stalls = []
        for h in hdr:
            m = re.match(r"smsp__average_warps?_issue_stalled_(\w+)_per_issue_active\.ratio", h) or \
                re.match(r"smsp__average_warp_latency_issue_stalled_(\w+)\.ratio", h)
            if m:
                stalls.append((num(r[col[h]]), m.group(1)))
        stalls = [s for s in stalls if s[0] == s[0]]
        stalls.sort(reverse=True)
        if stalls:
            lines.append("| top stall reasons (warps per issue-active cycle) | " +
                         ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]) + " |")
        lines.append("")
        rd = num(r[col["dram__bytes_read.sum"]]) * UNIT_SCALE.get(units[col["dram__bytes_read.sum"]], 1)
        wr = num(r[col["dram__bytes_write.sum"]]) * UNIT_SCALE.get(units[col["dram__bytes_write.sum"]], 1)
        base = re.sub(r"<.*$", "", short)
        if rd == rd and wr == wr:   # (a kernel ncu could not replay fully has NaN here)
            traffic[base] = int(rd + wr)
        k = "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"
        if k in col and num(r[col[k]]) == num(r[col[k]]):
            l1pct[base] = num(r[col[k]])
        # LSU data-pipe wavefronts of the launch = utilisation x elapsed L1 cycles summed over the SMs (the raw page carries
        # the percentage only); per warp-wide global load: 4 is the floor for 16-byte lanes (512 B / 128 B per wavefront)
        kc, kr, ks = "l1tex__cycles_elapsed.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"
        if k in col and kc in col and kr in col and num(r[col[kr]]) > 0 and num(r[col[k]]) == num(r[col[k]]):
            total = num(r[col[k]]) / 100.0 * num(r[col[kc]])
            shared = num(r[col[ks]]) if ks in col else 0.0
            wpr[base] = (total - shared) / num(r[col[kr]])
            lines.insert(len(lines) - 1, f"| LSU data-pipe wavefronts, derived (`{k}` x `{kc}`) | {total:.4g} |")
            lines.insert(len(lines) - 1, f"| of which shared memory | {shared:.4g} ({100 * shared / total:.0f} %) |")
            lines.insert(len(lines) - 1, f"| global-memory data-pipe wavefronts (loads + reds) per warp-wide global LOAD (floor 4 for 16-byte lanes) | {wpr[base]:.2f} |")
    os.makedirs(os.path.dirname(os.path.abspath(args.out_md)), exist_ok=True)
    with open(args.out_md, "w") as f:
        f.write("\n".join(lines))
    if args.traffic_key:
        tpath = os.path.join(os.path.dirname(os.path.abspath(args.out_md)), "traffic.json")
        data = {}
        if os.path.exists(tpath):
            with open(tpath) as f:
                data = json.load(f)
        data.setdefault(args.traffic_key, {}).update(traffic)
        data[args.traffic_key].setdefault("_l1_wavefront_pct", {}).update(l1pct)
        data[args.traffic_key].setdefault("_wavefronts_per_request", {}).update(wpr)
        data[args.traffic_key]["_source"] = os.path.basename(args.out_md)
        with open(tpath, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
