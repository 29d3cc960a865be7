#!/bin/bash
# Last refresh of the round: full GPU suite, smoke, default bench lines, codec launch lists (decode + encode), sanitizer on the codec.
TAG=${1:-kz}
O=gpurun_out
mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 100 --warmup 5 > $O/${TAG}_bench_config3.json 2> $O/${TAG}_bench_config3.err; echo "bench rc=$?"
timeout 400 python bench.py --workload config2 --steps 100 --warmup 5 --no-gpu-eager > $O/${TAG}_bench_config2.json 2> $O/${TAG}_bench_config2.err
CODEC_B=4 timeout 600 ncu --metrics $M --clock-control none --csv --log-file $O/${TAG}_launches_codec.csv python scripts/codec_probe.py 4545 > $O/${TAG}_ncu_codec.log 2>&1
python scripts/summarize_codec_launches.py $O/${TAG}_launches_codec.csv 45 > $O/${TAG}_launches_codec_summary.txt 2>&1
gzip -f $O/${TAG}_launches_codec.csv
timeout 600 ncu --metrics $M --clock-control none --csv --log-file $O/${TAG}_launches_codec_encode.csv python scripts/codec_encode_probe.py > $O/${TAG}_ncu_codec_enc.log 2>&1
python scripts/summarize_codec_launches.py $O/${TAG}_launches_codec_encode.csv 46 > $O/${TAG}_launches_codec_encode_summary.txt 2>&1
gzip -f $O/${TAG}_launches_codec_encode.csv
tail -3 $O/${TAG}_launches_codec_encode_summary.txt | cut -c1-250
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  timeout 900 $CS --tool $tool --print-limit 20 python scripts/sanitize_driver.py codec > $O/${TAG}_sanitize_${tool}_codec.log 2>&1
  echo "$tool codec: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/${TAG}_sanitize_${tool}_codec.log | tail -1)"
done
JEN1_LSTM=smem timeout 900 $CS --tool memcheck --print-limit 20 python scripts/sanitize_driver.py codec > $O/${TAG}_sanitize_memcheck_codec_smemlstm.log 2>&1
echo "memcheck codec (JEN1_LSTM=smem): $(grep -E 'ERROR SUMMARY' $O/${TAG}_sanitize_memcheck_codec_smemlstm.log | tail -1)"
python - <<PY
import json
for f in ("config3","config2"):
    d=json.loads(open("$O/${TAG}_bench_%s.json" % f).read().strip().splitlines()[-1])
    e=d["extra"]
    print(f, "ms/step %.4f" % d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "decode", e["codec_decode"]["ms_per_decode"], "encode", e["codec_encode"]["ms_per_encode"], "generate", (e.get("generate_audio") or {}).get("ms_per_call"))
PY
