#!/usr/bin/env python
"""Rank the SASS instructions of one kernel in an .ncu-rep by warp-stall samples (ncu --page source; needs
--import-source on at capture time).  Prints samples, executions and the instruction text, plus the unique executed
instruction footprint (what the 32 KB instruction cache has to hold).

    python scripts/ncu_lines.py <report.ncu-rep> [launch index = 0] [top = 30]
"""
import csv
import subprocess
import sys


def main(rep, launch="0", top="30"):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, sect, data = None, -1, []
    for r in rows:
        if len(r) > 5 and r[0] == "Address":
            hdr = r
            sect += 1
            continue
        if hdr and len(r) == len(hdr) and sect == int(launch):
            d = dict(zip(hdr, r))
            try:
                data.append((int(d["# Samples"]), int(d["Instructions Executed"]), d["Source"].strip()[:110]))
            except (KeyError, ValueError):
                pass
    tot = sum(d[0] for d in data)
    print("SASS instructions %d, executed at least once %d (%.1f KB footprint), warp-instructions executed %d, stall samples %d"
          % (len(data), sum(1 for d in data if d[1] > 0), sum(1 for d in data if d[1] > 0) * 16 / 1024.0, sum(d[1] for d in data), tot))
    for s, e, t in sorted(data, reverse=True)[: int(top)]:
        print("%6d %9d  %s" % (s, e, t))


if __name__ == "__main__":
    main(*sys.argv[1:])
