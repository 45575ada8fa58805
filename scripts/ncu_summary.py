#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) and an ncu launch list (csv) into small text files under profiles/.

    python scripts/ncu_summary.py report gpurun_out/prof.ncu-rep profiles/r1_stage3_full.txt
    python scripts/ncu_summary.py launches gpurun_out/launches.csv profiles/r1_launches.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_active.avg",
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"


def report(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        d = OrderedDict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append("== kernel: %s  (id %s)" % (d.get("Kernel Name", "?"), d.get("ID", "?")))
        for k in KEYS:
            if k in d and d[k] != "":
                lines.append("  %-82s %s %s" % (k, d[k], u.get(k, "")))
        stalls = [(float(v), k) for k, v in d.items() if k.startswith(STALL_PREFIX) and k.endswith("_per_issue_active.ratio") and v]
        for v, k in sorted(stalls, reverse=True)[:8]:
            lines.append("  stall %-76s %.3f" % (k[len(STALL_PREFIX):], v))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    iu = hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        t = float(r[iv].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        name = r[ik].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t * scale
    total = sum(a[1] for a in agg.values()) or 1.0
    lines = ["%-70s %8s %12s %7s" % ("kernel", "launches", "total ms", "share")]
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-70s %8d %12.3f %6.1f%%" % (name[:70], n, ms, 100 * ms / total))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    {"report": report, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
