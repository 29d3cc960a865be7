#!/usr/bin/env python
"""Rank source lines of one kernel in an .ncu-rep by stall samples / executed instructions (needs -lineinfo + --import-source)."""
import csv
import subprocess
import sys


def main(rep, kid="0", top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-id", "::regex:.*:" + kid],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur, data = None, []
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if len(r) > 8 and r[0].isdigit() and r[2] == '-':
            try:
                data.append((int(r[6]), int(r[7]), cur, int(r[0]), r[1].strip()[:120]))
            except ValueError:
                pass
    tot, toti = sum(d[0] for d in data), sum(d[1] for d in data)
    print('total samples', tot, 'warp-instructions', toti)
    for d in sorted(data, reverse=True)[:int(top)]:
        print("%6d %9d %s:%d: %s" % d)


if __name__ == "__main__":
    main(*sys.argv[1:])
