#!/bin/bash
# ncu evidence for profiles/: per-step launch lists (time + DRAM bytes) of bench.py for both workloads and full-set
# captures of the dominant kernels.  Usage (under gpurun): bash scripts/gpu_profile.sh <tag>
TAG=${1:-p}
O=gpurun_out
mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
# launch list: skip everything before the timed region's last 2 steps is fragile; capture all launches of a 3+2-step run
timeout 500 ncu --metrics $M --clock-control none -c 2200 --csv --log-file $O/${TAG}_launches_config2.csv python bench.py --workload config2 --steps 2 --warmup 3 --quick --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_c2.log 2>&1
timeout 500 ncu --metrics $M --clock-control none -c 2200 --csv --log-file $O/${TAG}_launches_config3.csv python bench.py --workload config3 --steps 2 --warmup 3 --quick --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_c3.log 2>&1
python scripts/summarize_launches.py $O/${TAG}_launches_config2.csv > $O/${TAG}_launches_config2_summary.txt 2>&1
python scripts/summarize_launches.py $O/${TAG}_launches_config3.csv > $O/${TAG}_launches_config3_summary.txt 2>&1
gzip -f $O/${TAG}_launches_config2.csv $O/${TAG}_launches_config3.csv
# full captures (second, warm UNet evaluation of scripts/timeline.py --plain): a deep split-K conv, a hi-res conv, attention
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 315 -c 2 -o $O/${TAG}_full_conv_deep_c2 -f python scripts/timeline.py --plain 1515 1 > $O/${TAG}_full1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 237 -c 2 -o $O/${TAG}_full_conv_hires_c3 -f python scripts/timeline.py --plain 4545 4 > $O/${TAG}_full2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_umma -s 26 -c 2 -o $O/${TAG}_full_attn_c3 -f python scripts/timeline.py --plain 4545 4 > $O/${TAG}_full3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_flash -c 2 -o $O/${TAG}_full_attn_flash -f python scripts/attn_bench.py --once --shapes 4545,8,128,0 4545,8,64,0 > $O/${TAG}_full4.log 2>&1
# Encodec decoder: launch list of one 30 s decode (4 samples) + full captures of its three kernel classes
CODEC_B=4 timeout 600 ncu --metrics $M --clock-control none --csv --log-file $O/${TAG}_launches_codec.csv python scripts/codec_probe.py 4545 > $O/${TAG}_ncu_codec.log 2>&1
python scripts/summarize_codec_launches.py $O/${TAG}_launches_codec.csv > $O/${TAG}_launches_codec_summary.txt 2>&1
gzip -f $O/${TAG}_launches_codec.csv
CODEC_B=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tf32 -s 30 -c 2 -o $O/${TAG}_full_codec_conv_tf32 -f python scripts/codec_probe.py 4545 > $O/${TAG}_full5.log 2>&1
CODEC_B=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:lstm_tc -c 1 -o $O/${TAG}_full_codec_lstm -f python scripts/codec_probe.py 600 > $O/${TAG}_full6.log 2>&1
CODEC_B=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:narrow_conv -c 1 -o $O/${TAG}_full_codec_narrow -f python scripts/codec_probe.py 4545 > $O/${TAG}_full7.log 2>&1
tail -3 $O/${TAG}_launches_codec_summary.txt
head -12 $O/${TAG}_launches_config2_summary.txt; head -8 $O/${TAG}_launches_config3_summary.txt
