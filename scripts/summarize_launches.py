#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name (share of the step)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot, n = 0.0, 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        agg[k][0] += 1
        agg[k][1] += v
        tot += v
        n += 1
    print("%d launches, %.1f us total (cold-cache, serialised: compare shares)" % (n, tot))
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-78s %6d %12.1f us %6.2f%%" % (k[:78], c, t, 100 * t / tot))


if __name__ == "__main__":
    main(sys.argv[1])
