#!/bin/bash
# One GPU-box pass: parity tests, bench lines, launch list, timeline, one full ncu capture of the tcgen05 conv kernel.
# Usage (under gpurun): bash scripts/gpu_check.sh <tag>
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 5 > $O/${TAG}_bench_config2.json 2> $O/${TAG}_bench_config2.err
timeout 600 python bench.py --workload config3 --steps 30 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_config3.json 2> $O/${TAG}_bench_config3.err
JEN1_TIMELINE=1 timeout 300 python scripts/timeline.py 1515 1 > /dev/null 2> $O/${TAG}_timeline_c2.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${TAG}_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
python scripts/summarize_launches.py $O/${TAG}_launches_c2.csv > $O/${TAG}_launches_c2_summary.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 300 -c 24 -o $O/${TAG}_umma_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_full.log 2>&1
gzip -f $O/${TAG}_launches_c2.csv
tail -3 $O/${TAG}_pytest_gpu.log; cat $O/${TAG}_bench_config2.json; cat $O/${TAG}_bench_config3.json; cat $O/${TAG}_launches_c2_summary.txt | head -20
