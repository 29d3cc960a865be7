#!/bin/bash
# Encodec decoder pass: parity check, timing, ncu launch list.  Usage (under gpurun): bash scripts/gpu_codec.sh <tag>
TAG=${1:-k}
O=gpurun_out
mkdir -p $O
timeout 600 python scripts/codec_check.py 4545 4 2>&1 | tail -12
timeout 300 python scripts/codec_probe.py 150 1515 4545 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/${TAG}_codec_launches.csv python scripts/codec_probe.py 4545 > $O/${TAG}_codec_ncu.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$O/${TAG}_codec_launches.csv", errors="replace")) if len(r)>14 and r[0].isdigit()]
L={}
for r in rows:
    d=L.setdefault(int(r[0]), {"name": r[4][:60], "grid": r[8]})
    d[r[12]]=float(r[14].replace(",",""))
ids=sorted(L)
half=ids[len(ids)//2:]   # the second (warm) decode
tot=sum(L[i].get("gpu__time_duration.sum",0) for i in half)
print("second decode: %d launches, %.2f ms serialised" % (len(half), tot/1e6))
for i in half:
    d=L[i]; print("%4d %-60s %-16s %9.1f us  rd %8.1f MB wr %8.1f MB" % (i, d["name"], d["grid"], d.get("gpu__time_duration.sum",0)/1e3, d.get("dram__bytes_read.sum",0)/1e6, d.get("dram__bytes_write.sum",0)/1e6))
PY
