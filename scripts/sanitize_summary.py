#!/usr/bin/env python
"""Digest of a scripts/sanitize.sh run: python scripts/sanitize_summary.py <tag> > profiles/rNN_sanitize.txt"""
import glob
import re
import sys


def main(tag):
    print("# compute-sanitizer pass (scripts/sanitize.sh %s) over scripts/sanitize_driver.py: 'tiny' = tiny UNet, fp32 generic +" % tag)
    print("# bf16 tcgen05 engines, one forward + a 3-step sampler; 'full' = full model at T=47, bf16 tcgen05 engine (every kernel of the benchmarked path)")
    for f in sorted(glob.glob("gpurun_out/%s_sanitize_*.log" % tag)):
        txt = open(f, errors="replace").read()
        summ = re.findall(r"(ERROR SUMMARY: .*|RACECHECK SUMMARY: .*)", txt)
        kinds = {}
        for m in re.finditer(r"=========\s+(Invalid [^\n]*|Uninitialized [^\n]*|Barrier error[^\n]*|(?:Error|Warning): (?:Race|Potential)[^\n]*)", txt):
            k = re.sub(r"0x[0-9a-f]+", "0x..", m.group(1))[:110]
            kinds[k] = kinds.get(k, 0) + 1
        ok = re.findall(r"^((?:tiny|full|codec) [^\n]*ok[^\n]*)$", txt, re.M)
        print("%-46s %s" % (f.split("/")[-1], summ[-1] if summ else "NO SUMMARY (timeout?)"))
        for k, n in sorted(kinds.items(), key=lambda kv: -kv[1])[:6]:
            print("      %5d x %s" % (n, k))
        for o in ok[-2:]:
            print("      " + o)
    for f in sorted(glob.glob("gpurun_out/%s_stress.log" % tag)):
        print("%-46s %s" % (f.split("/")[-1], open(f).read().strip().splitlines()[-1]))


if __name__ == "__main__":
    main(sys.argv[1])
