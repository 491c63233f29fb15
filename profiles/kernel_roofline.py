#!/usr/bin/env python
"""Per-kernel roofline table of one FM-index build (BASELINE.json north_star: "every kernel's achieved HBM GB/s
reported against peak").

  python profiles/kernel_roofline.py <launches.csv> <stats.json> <out.json>

launches.csv: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none
--csv` of a run that builds the workload twice (tools_dev/one_build.py C3 2): the SECOND build is taken (warm
allocations).  Times under ncu are cold-cache and serialised: the table is for SHARES and for traffic, the bench's
own CUDA-event timings are the numbers of record.  stats.json: {"n": symbols, "members": refine_members,
"keys": refine_key_fetches, "section_bytes": bytes of [data|Rs|Rb]} of that build.

Per kernel: launches, ms, algorithmic bytes (the model below: what the kernel must read and write at least),
DRAM bytes ncu counted, achieved = algorithmic bytes / time, frac = achieved / measured HBM peak
(MEASURED_PEAKS.json hbm_gbs), traffic ratio = DRAM bytes / algorithmic bytes.
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def model(n, members, keys, section_bytes):
    """kernel name prefix -> (algorithmic bytes over ALL its launches of one build, what is counted)"""
    data_bytes = section_bytes * 253 / 347.0  # share of the bit words in [data | Rs | Rb] (8 : 2 : 1 per 64 bits)
    return collections.OrderedDict([
        ("byte_hist_kernel", (n, "raw text read")),
        ("doc_stats_kernel", (n, "raw text read")),
        ("pack_kernel", (n + 3 * n / 8, "raw text read, 3-bit packed text written")),
        ("make_keys_hist_kernel", (3 * n / 8 + 8 * n, "packed text read, 8-byte keys written")),
        ("radix_scan_kernel", (6 * 256 * 16, "digit tables")),
        ("onesweep_kernel<1", (8 * n + 12 * n, "first pass: keys read (positions are the indices), keys + positions written")),
        ("onesweep_kernel<0", (5 * 24 * n, "five passes: 8+4 bytes read and written per pair")),
        ("heads_kernel", (8 * n + n + n / 4, "sorted keys read; BWT bytes and two bitmaps written")),
        ("mark_active_kernel", (3 * n / 8, "two bitmaps read, one written")),
        ("count_active_kernel", (n / 8, "bitmap read")),
        ("scan_tiles_kernel", (n / 8192 * 16, "tile counts")),
        ("compact_active_kernel", (n / 4 + members * (5 + 9), "bitmaps; suffix + BWT byte of every member read; suffix, slot, byte written")),
        ("pad_heads_kernel", (64, "tail of a bitmap")),
        ("refine_warps_kernel", (members * 5 * 2 + keys * 8, "suffix + byte of every member read and written back; 8 bytes per key gathered")),
        ("scatter_bwt_kernel", (members * 5 + members, "slot + byte read, byte written")),
        ("wt_sweep_kernel", (n + data_bytes, "BWT read, node bit words written")),
        ("rank_chunk_sum_kernel", (data_bytes, "bit words read")),
        ("rank_chunk_scan_kernel", (6 * 8 * (n / 256 / 2048 + 1), "chunk sums")),
        ("rank_write_kernel", (section_bytes, "bit words read, Rs and Rb written")),
    ])


def main():
    src, stats_path, dst = sys.argv[1:4]
    st = json.load(open(stats_path))
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    lines = [l for l in open(src) if not l.startswith("==")]
    by = collections.OrderedDict()
    for r in csv.DictReader(lines):
        by.setdefault(r["ID"], {"name": r["Kernel Name"]})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    seq = list(by.values())
    starts = [i for i, x in enumerate(seq) if "byte_hist" in x["name"]]
    build = seq[starts[-1]:]
    mdl = model(st["n"], st["members"], st["keys"], st["section_bytes"])
    agg = collections.OrderedDict()
    for x in build:
        name = x["name"].split("(")[0].replace("void ", "").replace("dsmfm::<unnamed>::", "").replace("unnamed>::", "")
        key = next((k for k in mdl if name.startswith(k)), name)
        a = agg.setdefault(key, {"kernel": name, "launches": 0, "ms": 0.0, "dram_bytes": 0.0})
        a["launches"] += 1
        a["ms"] += x["gpu__time_duration.sum"] / 1e6
        a["dram_bytes"] += x.get("dram__bytes_read.sum", 0.0) + x.get("dram__bytes_write.sum", 0.0)
    total = sum(a["ms"] for a in agg.values())
    rows = []
    for key, a in agg.items():
        alg, what = mdl.get(key, (None, "no model"))
        row = {"kernel": a["kernel"], "launches": a["launches"], "ms": round(a["ms"], 3), "share": round(a["ms"] / total, 4),
               "algorithmic_bytes": None if alg is None else int(alg), "counted": what, "dram_bytes": int(a["dram_bytes"])}
        if alg:
            ach = alg / (a["ms"] * 1e-3) / 1e9
            row.update({"achieved_gbs": round(ach, 1), "frac_of_peak": round(ach / peak, 4),
                        "dram_over_algorithmic": round(a["dram_bytes"] / alg, 2)})
        rows.append(row)
    out = {"workload": st.get("workload"), "n": st["n"], "peak_gbs": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)",
           "total_ms_under_ncu": round(total, 3), "note": "times under ncu (cold cache, serialised): shares and traffic, not bench values",
           "kernels": rows}
    json.dump(out, open(dst, "w"), indent=1)
    for r in rows:
        print("%-34s %3d %8.3f ms %5.1f%%  alg %8.3f GB  dram %8.3f GB  %7s GB/s  frac %s" % (
            r["kernel"][:34], r["launches"], r["ms"], 100 * r["share"], (r["algorithmic_bytes"] or 0) / 1e9, r["dram_bytes"] / 1e9,
            r.get("achieved_gbs", "-"), r.get("frac_of_peak", "-")))


if __name__ == "__main__":
    main()
