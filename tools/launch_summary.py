#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: total, calls, average, share.
Usage: tools/launch_summary.py gpurun_out/launches_final.csv"""
import collections, csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(r[ui], 1e-6)
    name = re.sub(r"\(.*", "", r[ki])[:75]
    tot[name] += v; cnt[name] += 1
s = sum(tot.values())
print("  total ms  calls    avg us   share  kernel")
for k, v in tot.most_common(24):
    print("  %8.3f  %5d  %8.1f  %5.1f%%  %s" % (v, cnt[k], v / cnt[k] * 1e3, 100 * v / s, k))
