"""Summarise an ncu gpu__time_duration launch list (csv) per kernel: count, avg us, share."""
import collections
import csv
import sys

for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        k = r[ki]
        if k.startswith("void dmr::"):
            k = k[5:]
        k = k[:k.index("(")] if "(" in k and k.startswith("dmr::") else k[:70]
        agg.setdefault(k, []).append(float(r[vi].replace(",", "")))
    ours = {k: v for k, v in agg.items() if k.startswith("dmr::")}
    tot = sum(sum(v) for v in ours.values())
    tot_all = sum(sum(v) for v in agg.values())
    print("== %s  (our kernels %.1f us of %.1f us total)" % (path, tot / 1000, tot_all / 1000))
    for k, v in ours.items():
        print("  %-38s n=%3d avg %10.1f us  share %5.1f%%" % (k, len(v), sum(v) / len(v) / 1000, 100 * sum(v) / tot))
