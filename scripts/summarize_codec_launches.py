#!/usr/bin/env python
"""Per-launch table of one decode from an ncu launch list of scripts/codec_probe.py (last decode in the file)."""
import csv
import sys


def main(path, launches_per_decode=24):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14 and r[0].isdigit()]
    L = {}
    for r in rows:
        d = L.setdefault(int(r[0]), {"name": r[4].replace("jen1::", "").replace("<unnamed>::", "")[:58], "grid": r[8]})
        d[r[12]] = float(r[14].replace(",", ""))
    ids = sorted(L)[-launches_per_decode:]
    tot = sum(L[i].get("gpu__time_duration.sum", 0) for i in ids)
    print("# last decode of the run: %d launches, %.2f ms serialised under ncu (cold cache)" % (len(ids), tot / 1e6))
    agg = {}
    for i in ids:
        d = L[i]
        k = d["name"].split("(")[0]
        agg[k] = agg.get(k, 0.0) + d.get("gpu__time_duration.sum", 0)
        print("%4d %-58s %-16s %9.1f us  rd %8.1f MB  wr %8.1f MB" % (i, d["name"], d["grid"], d.get("gpu__time_duration.sum", 0) / 1e3,
              d.get("dram__bytes_read.sum", 0) / 1e6, d.get("dram__bytes_write.sum", 0) / 1e6))
    print("# share by kernel:", ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 24)
