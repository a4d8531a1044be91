#!/usr/bin/env python
"""Turn the raw ncu output of a gpurun session (gpurun_out/, scratch) into the text summaries under profiles/.

  launches <csv> <out.txt> <title>                    per-kernel launch list (ncu --metrics gpu__time_duration.sum --csv)
  kernel   <rep> <out.txt> <title> [agent_steps]      selected metrics of every launch in an ncu --set full report
  traffic  <out.json> name=<rep> ...                  dram bytes (read + write) of the first launch of each report
"""
import collections
import csv
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def to_bytes(val, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(val) * scale


def cmd_launches(path, out, title):
    agg = collections.OrderedDict()
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    for r in rows[1:]:
        if r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
            continue
        us = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
        a = agg.setdefault(r[ik], [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n# (per-launch times under ncu are cold-cache / serialised: compare shares, not absolutes)\n")
        f.write(f"{'kernel':90s} {'launches':>8s} {'total_us':>12s} {'avg_us':>9s} {'share':>7s}\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:90]:90s} {n:8d} {us:12.1f} {us / n:9.2f} {us / tot * 100:6.1f}%\n")
    print(open(out).read())


def cmd_kernel(rep, out, title, agent_steps=None, bytes_per_unit=None):
    hdr, units, rows = raw_rows(rep)
    ik = hdr.index("Kernel Name")
    with open(out, "w") as f:
        f.write(f"# {title}\n")
        for n, r in enumerate(rows):
            f.write(f"## launch {n}: {r[ik][:80]}\n")
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"{m:85s} {units[i]:14s} {r[i]}\n")
            rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
            wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
            line = f"dram traffic per launch: {rd + wr:.0f} B"
            if agent_steps:
                line += f" = {(rd + wr) / float(agent_steps):.1f} B per unit (algorithmic {bytes_per_unit or 336})"
            f.write(line + "\n")
    print(open(out).read()[:3000])


def cmd_traffic(out, pairs):
    res = {}
    try:
        res = json.load(open(out))
    except Exception:
        pass
    for p in pairs:
        name, rep = p.split("=", 1)
        hdr, units, rows = raw_rows(rep)
        r = rows[0]
        res[name] = sum(to_bytes(r[hdr.index(m)], units[hdr.index(m)]) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    json.dump(res, open(out, "w"), indent=1)
    print(res)


if __name__ == "__main__":
    c = sys.argv[1]
    if c == "launches":
        cmd_launches(*sys.argv[2:5])
    elif c == "kernel":
        cmd_kernel(*sys.argv[2:7])
    elif c == "traffic":
        cmd_traffic(sys.argv[2], sys.argv[3:])
