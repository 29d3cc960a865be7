#!/usr/bin/env python
"""Executed instruction footprint of conv_umma_kernel per source line: joins the SASS page of an .ncu-rep (which
instructions executed at least once in one launch) with nvdisasm's line info of the in-tree object.

    python scripts/ncu_footprint.py <report.ncu-rep> [launch index = 0] [object = jen1_b200/_C/conv_umma.o]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile


def main(rep, launch="0", obj="jen1_b200/_C/conv_umma.o"):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, sect, ex = None, -1, []
    for r in rows:
        if len(r) > 5 and r[0] == "Address":
            hdr, sect = r, sect + 1
            continue
        if hdr and len(r) == len(hdr) and sect == int(launch):
            d = dict(zip(hdr, r))
            ex.append((int(d["Instructions Executed"]), int(d["# Samples"]), d["Source"].strip()))
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    line, inl, instrs = 0, "", []
    for ln in dis.splitlines():
        m = re.search(r'//## File ".*?([^/"]+)", line (\d+)(.*)', ln)
        if m:
            line, inl = int(m.group(2)), m.group(1)
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            instrs.append((inl, line, m.group(2)))
    print("ncu rows %d, nvdisasm instructions %d" % (len(ex), len(instrs)))
    n = min(len(ex), len(instrs))
    per = {}
    for i in range(n):
        k = instrs[i][:2]
        a = per.setdefault(k, [0, 0, 0, 0])
        a[0] += 1
        a[1] += 1 if ex[i][0] > 0 else 0
        a[2] += ex[i][0]
        a[3] += ex[i][1]
    print("total %d, executed once %d" % (n, sum(a[1] for a in per.values())))
    print("file:line   sass  executed-once  warp-instr  samples")
    for k, a in sorted(per.items(), key=lambda kv: -kv[1][1])[:70]:
        print("%s:%d  %5d %5d %9d %5d" % (k[0], k[1], a[0], a[1], a[2], a[3]))


if __name__ == "__main__":
    main(*sys.argv[1:])
