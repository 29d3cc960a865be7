#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list per
kernel name: launches, device time and share of the total, DRAM bytes.  Times are cold-cache and serialised (ncu
replays every kernel alone): compare SHARES, not absolutes.

    python scripts/summarize_launches.py launches.csv[.gz] [steps]   # steps: divide totals to get per-step figures
"""
import collections
import csv
import gzip
import re
import sys


def to_us(v, u):
    return v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else (v * 1e6 if u in ("s", "second") else v))


def to_bytes(v, u):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return v * m.get(u, 1)


def main(path, steps=1):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for row in csv.DictReader(lines):
        k = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        n, u = row.get("Metric Name"), row["Metric Unit"]
        a = agg[k]
        if n == "gpu__time_duration.sum":
            a[0] += 1
            a[1] += to_us(v, u)
        elif n == "dram__bytes_read.sum":
            a[2] += to_bytes(v, u)
        elif n == "dram__bytes_write.sum":
            a[3] += to_bytes(v, u)
    tot = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    rd, wr = sum(a[2] for a in agg.values()), sum(a[3] for a in agg.values())
    print("%d launches, %.1f us total (cold-cache, serialised: compare shares), DRAM read %.1f MB write %.1f MB" % (n, tot, rd / 1e6, wr / 1e6))
    if steps > 1:
        print("per step (/%d): %.1f launches, %.1f us, DRAM read %.1f MB write %.1f MB" % (steps, n / steps, tot / steps, rd / steps / 1e6, wr / steps / 1e6))
    print("%-64s %7s %12s %7s %10s %10s" % ("kernel", "count", "time us", "share", "rd MB", "wr MB"))
    for k, (c, t, r, w) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-64s %7d %12.1f %6.2f%% %10.1f %10.1f" % (k[:64], c, t, 100 * t / tot, r / 1e6, w / 1e6))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
