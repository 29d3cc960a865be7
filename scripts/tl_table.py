#!/usr/bin/env python
"""Turn a JEN1_TIMELINE=1 JEN1_TRACE=1 log (scripts/timeline.py) into a per-op table (us) of the warm evaluation."""
import re
import sys

F = 1.965e3  # SM clock (MHz -> cycles per us)


def main(path):
    lines = open(path).read().split('==== second (warm) evaluation')[1].splitlines()
    ops = [l for l in lines if l.startswith('[jen1] op') and 'umma ' in l]
    tls = [l for l in lines if l.startswith('[jen1-tl] u')]
    tot = {}
    for o, t in zip(ops, tls):
        m = re.search(r'B=(\d+) Lm=(\d+) Lout=(\d+) Cin=(\d+)\(\+(\d+)\) Cout=(\d+) taps=(\d+) stride=(\d+) phases=(\d+) G=(\d+) mode=(\d+) \| NT=(\d+) tiles=(\d+)x(\d+) splitk=(\d+) stages=(\d+)', o)
        B, Lm, Lo, Ci, Ci2, Co, taps, st, ph, G, mode, NT, nt, mt, sk, stg = map(int, m.groups())
        d = dict(re.findall(r'(\w+) (-?\d+)', t.split('|', 1)[0]))
        gap, body = int(d['gap_ns']) / 1e3, int(d['body_ns']) / 1e3
        c = dict(re.findall(r'(\w+) (-?\d+)', t.split('cyc:')[1]))
        g = lambda k: int(c[k]) / F
        print("%s B%d L%4d Ci%4d+%4d Co%4d k%d s%d p%d G%2d m%d NT%3d t%3dx%d sk%2d st%d | gap %5.1f body %5.1f | early %6.1f stats %4.1f coef %4.1f panels %5.1f acc %5.1f part %5.1f clus %5.1f eploop %5.1f epstats %5.1f end %5.1f | first_a %4.1f issued %4.1f"
              % (o[7:12], B, Lm, Ci, Ci2, Co, taps, st, ph, G, mode, NT, nt, mt, sk, stg, gap, body, g('early'), g('stats'), g('coef'),
                 g('panels'), g('accfull'), g('part') if 'part' in c else 0.0, g('cluster'), g('eploop') if 'eploop' in c else 0.0,
                 g('epstats') if 'epstats' in c else 0.0, g('end'), g('first_a'), g('issued')))
        key = "L%d" % Lm
        a = tot.setdefault(key, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += gap
        a[2] += body
    print("--- per output length: ops, sum gap us, sum body us")
    for k, (n, ga, bo) in tot.items():
        print("%-8s %3d %8.1f %8.1f" % (k, n, ga, bo))


if __name__ == "__main__":
    main(sys.argv[1])
