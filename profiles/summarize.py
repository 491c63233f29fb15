#!/usr/bin/env python
"""Summarise ncu artefacts brought back in gpurun_out/ into small text files under profiles/.
  python profiles/summarize.py launches gpurun_out/launches.csv profiles/rN_launches.txt
  python profiles/summarize.py full gpurun_out/prof.ncu-rep profiles/rN_kernel_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|dram__sectors_(read|write)\.sum|"
                  r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__warps_active\.avg\.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|"
                  r"launch__block_size|launch__occupancy_limit_\w+|launch__shared_mem_per_block_\w+|smsp__inst_executed\.sum|"
                  r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__thread_inst_executed_per_inst_executed\.ratio|"
                  r"lts__t_sector_hit_rate\.pct|l1tex__t_sector_hit_rate\.pct|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|"
                  r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio)$")


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").replace("dsmfm::<unnamed>::", "").replace("unnamed>::", "").replace("dsmfm::<", "")


def launches(src, dst):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    seq = [(short(r["Kernel Name"]), float(r["Metric Value"].replace(",", "")) / 1e6, r["Grid Size"]) for r in csv.DictReader(lines)]
    starts = [i for i, (n, _, _) in enumerate(seq) if "byte_hist" in n]
    build = seq[starts[0]:starts[1]] if len(starts) > 1 else seq
    tot = sum(m for _, m, _ in build)
    agg = collections.OrderedDict()
    for n, m, g in build:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += m
    with open(dst, "w") as out:
        out.write("# one FM-index build under `ncu --metrics gpu__time_duration.sum --clock-control none`\n")
        out.write("# (cold-cache, serialised launches: compare SHARES, not absolutes)\n")
        out.write("%-34s %5s %10s %7s\n" % ("kernel", "calls", "ms", "share"))
        for n, (c, m) in agg.items():
            out.write("%-34s %5d %10.3f %6.1f%%\n" % (n, c, m, 100 * m / tot))
        out.write("%-34s %5d %10.3f\n\n# every launch in order: kernel, ms, grid\n" % ("total", len(build), tot))
        for n, m, g in build:
            out.write("%s\t%.3f\t%s\n" % (n, m, g))


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as out:
        out.write("# ncu --set full --clock-control none, %d launch(es) of %s\n" % (len(rows) - 2, short(rows[2][hdr.index("Kernel Name")])))
        for i, h in enumerate(hdr):
            if KEEP.match(h):
                out.write("%-75s %-14s %s\n" % (h, units[i], "  ".join(r[i] for r in rows[2:])))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
