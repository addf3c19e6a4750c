#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum[,smsp__inst_executed.sum] --csv):
total time, launches and (if present) warp instructions per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
kn, mn, mv, idc = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
tot = collections.Counter()
ins = collections.Counter()
ids = collections.defaultdict(set)
for r in rows[1:]:
    name = r[kn].split("(")[0].replace("void ", "").replace("vgc::", "")
    v = float(r[mv].replace(",", ""))
    if r[mn].startswith("gpu__time_duration"):
        tot[name] += v
        ids[name].add(r[idc])
    elif r[mn].startswith("smsp__inst_executed"):
        ins[name] += v
all_ns = sum(tot.values())
print("%-24s %8s %11s %7s %9s %12s %10s" % ("kernel", "launches", "total ms", "share", "avg us", "Ginst", "inst/ns"))
for k, v in tot.most_common():
    n = len(ids[k])
    print("%-24s %8d %11.3f %6.1f%% %9.1f %12.3f %10.1f" % (k, n, v / 1e6, 100 * v / all_ns, v / n / 1e3, ins[k] / 1e9,
                                                        ins[k] / v if v else 0))
print("%-24s %8d %11.3f %26s %12.3f" % ("all", sum(len(x) for x in ids.values()), all_ns / 1e6, "", sum(ins.values()) / 1e9))
