#!/bin/bash
# Refresh of the Encodec-decoder evidence + the default bench line.  Usage (under gpurun): bash scripts/gpu_codec_final.sh <tag>
TAG=${1:-kf}
O=gpurun_out
mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -2; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
CODEC_B=4 CODEC_REPS=10 timeout 300 python scripts/codec_probe.py 4545 2>&1 | tail -2
unset JEN1_B200_LIB
timeout 900 python bench.py --steps 100 --warmup 5 > $O/${TAG}_bench_config3.json 2> $O/${TAG}_bench_config3.err; echo "bench rc=$?"
timeout 300 python bench.py --workload config2 --steps 100 --warmup 5 --no-gpu-eager > $O/${TAG}_bench_config2.json 2> $O/${TAG}_bench_config2.err
CODEC_B=4 timeout 600 ncu --metrics $M --clock-control none --csv --log-file $O/${TAG}_launches_codec.csv python scripts/codec_probe.py 4545 > $O/${TAG}_ncu_codec.log 2>&1
python scripts/summarize_codec_launches.py $O/${TAG}_launches_codec.csv 45 > $O/${TAG}_launches_codec_summary.txt 2>&1
gzip -f $O/${TAG}_launches_codec.csv
CODEC_B=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tf32 -s 37 -c 2 -o $O/${TAG}_full_codec_conv_tf32 -f python scripts/codec_probe.py 4545 > $O/${TAG}_full5.log 2>&1
CODEC_B=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:lstm_tc -c 1 -o $O/${TAG}_full_codec_lstm -f python scripts/codec_probe.py 600 > $O/${TAG}_full6.log 2>&1
CODEC_B=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:narrow_conv -c 1 -o $O/${TAG}_full_codec_narrow -f python scripts/codec_probe.py 4545 > $O/${TAG}_full7.log 2>&1
tail -4 $O/${TAG}_launches_codec_summary.txt
python - <<PY
import json
for f in ("config3","config2"):
    d=json.loads(open("$O/${TAG}_bench_%s.json" % f).read().strip().splitlines()[-1])
    print(f, "ms/step %.4f" % d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "codec", (d["extra"].get("codec_decode") or {}).get("ms_per_decode"))
PY
