#!/usr/bin/env python
"""Summarise an ncu launch list (gpu__time_duration.sum CSV) and a --set full report into text for profiles/."""
import collections
import csv
import subprocess
import sys


def launch_summary(path, skip_torch=False):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] in ("ns", "nsecond") else (v * 1000 if r[ui] in ("ms", "msecond") else v)
        agg.setdefault(r[ki], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    out = [f"total {tot:.1f} us over {sum(len(v) for v in agg.values())} launches (cold-cache, serialised: compare shares)"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"{100 * sum(v) / tot:6.2f}%  {sum(v):10.1f} us  {len(v):4d}x  avg {sum(v) / len(v):9.1f} us  {k[:110]}")
    return "\n".join(out)


METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
           "sm__inst_executed_pipe_tensor.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]


def full_summary(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        out.append(f"--- {d.get('Kernel Name', '?')[:100]}  (id {d.get('ID')})")
        for m in METRICS:
            if m in d:
                out.append(f"    {m:70s} {d[m]:>18s} {units[hdr.index(m)]}")
    return "\n".join(out)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        print(launch_summary(sys.argv[2]))
    else:
        print(full_summary(sys.argv[2]))
