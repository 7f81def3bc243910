#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name."""
import collections
import csv
import sys


def main(path, steps=1):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1e-3)
        a = agg.setdefault(row["Kernel Name"][:70], [0, 0.0])
        a[0] += 1; a[1] += v; n += 1
    tot = sum(a[1] for a in agg.values())
    print("launches %d  total %.1f ms  (%d steps captured -> %.1f ms/step)" % (n, tot / 1e3, steps, tot / 1e3 / steps))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:22]:
        print("%-72s n=%4d %9.2f ms/step %5.1f%%" % (k, a[0], a[1] / 1e3 / steps, 100 * a[1] / tot))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
