#!/usr/bin/env python
"""ncu_summary.py <ncu --csv log> <sites> -- one line per distinct kernel: launches, mean duration, DRAM bytes per site,
FP64-pipe %, LSU %, registers, achieved DRAM GB/s (the per-kernel table of profiles/r2_kernels.md)."""
import csv
import collections
import sys

path, sites = sys.argv[1], float(sys.argv[2])
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[1:]:
    if r[ix["ID"]] == "ID":
        continue
    key = (r[ix["ID"]], r[ix["Kernel Name"]])
    per.setdefault(key, {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
agg = collections.OrderedDict()
for (_, name), m in per.items():
    short = name.split("(")[0].replace("void ", "").replace("gfb::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    agg.setdefault(short, []).append(m)
print("| kernel | launches | mean us | DRAM B/site (r+w) | DRAM GB/s | fp64 pipe % | LSU % | regs | warps active % |")
print("|---|---|---|---|---|---|---|---|---|")
for name, ms in agg.items():
    n = len(ms)
    g = lambda k: sum(m.get(k, 0.0) for m in ms) / n
    t = g("gpu__time_duration.sum")
    by = g("dram__bytes_read.sum") + g("dram__bytes_write.sum")
    print("| `%s` | %d | %.1f | %.0f | %.0f | %.1f | %.1f | %d | %.1f |" % (
        name[:70], n, t / 1e3, by / sites, by / t if t else 0, g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        g("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"), g("launch__registers_per_thread"),
        g("sm__warps_active.avg.pct_of_peak_sustained_active")))
