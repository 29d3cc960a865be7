#!/bin/bash
# Round-end evidence pass (UNet path + attention; the codec kernels are covered by gpu_codec_final.sh): GPU tests, smoke, the
# default bench + other workloads, fused-transformer arm, timelines, attention evidence, ncu launch lists / full captures.
# Usage (under gpurun): bash scripts/gpu_final2.sh <tag>
TAG=${1:-fin}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -s --timeout 600 > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
grep -E "passed|failed|rc=" $O/${TAG}_pytest_gpu.log | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -3 $O/${TAG}_smoke.log
timeout 900 python bench.py --steps 100 --warmup 5 > $O/${TAG}_bench_config3.json 2> $O/${TAG}_bench_config3.err; echo "bench rc=$?"
timeout 300 python bench.py --workload config2 --steps 100 --warmup 5 --no-gpu-eager > $O/${TAG}_bench_config2.json 2> $O/${TAG}_bench_config2.err
timeout 300 python bench.py --workload config5 --steps 100 --warmup 5 --quick --no-cpu-baseline > $O/${TAG}_bench_config5.json 2> $O/${TAG}_bench_config5.err
timeout 300 python bench.py --scaling strong --steps 20 --warmup 3 --quick --no-cpu-baseline > $O/${TAG}_bench_strong_n1.json 2> $O/${TAG}_bench_strong_n1.err
JEN1_FUSED_TR=1 timeout 300 python bench.py --steps 50 --warmup 5 --quick --no-cpu-baseline --no-e2e > $O/${TAG}_bench_config3_fusedtr.json 2> /dev/null
JEN1_FUSED_TR=1 timeout 300 python bench.py --workload config2 --steps 50 --warmup 5 --quick --no-cpu-baseline --no-e2e > $O/${TAG}_bench_config2_fusedtr.json 2> /dev/null
JEN1_TIMELINE=1 timeout 120 python scripts/timeline.py 1515 1 > /dev/null 2> $O/${TAG}_timeline_c2.raw
JEN1_TIMELINE=1 timeout 120 python scripts/timeline.py 4545 4 > /dev/null 2> $O/${TAG}_timeline_c3.raw
python scripts/tl_table.py $O/${TAG}_timeline_c2.raw > $O/${TAG}_tl_c2.txt 2>&1
python scripts/tl_table.py $O/${TAG}_timeline_c3.raw > $O/${TAG}_tl_c3.txt 2>&1
timeout 300 python scripts/attn_bench.py > $O/${TAG}_attn_bench.jsonl 2> $O/${TAG}_attn_bench.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 400 ncu --metrics $M --clock-control none -c 2200 --csv --log-file $O/${TAG}_launches_config2.csv python bench.py --workload config2 --steps 2 --warmup 3 --quick --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_c2.log 2>&1
timeout 400 ncu --metrics $M --clock-control none -c 2200 --csv --log-file $O/${TAG}_launches_config3.csv python bench.py --workload config3 --steps 2 --warmup 3 --quick --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_c3.log 2>&1
python scripts/summarize_launches.py $O/${TAG}_launches_config2.csv > $O/${TAG}_launches_config2_summary.txt 2>&1
python scripts/summarize_launches.py $O/${TAG}_launches_config3.csv > $O/${TAG}_launches_config3_summary.txt 2>&1
gzip -f $O/${TAG}_launches_config2.csv $O/${TAG}_launches_config3.csv
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 315 -c 2 -o $O/${TAG}_full_conv_deep_c2 -f python scripts/timeline.py --plain 1515 1 > $O/${TAG}_full1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 237 -c 2 -o $O/${TAG}_full_conv_hires_c3 -f python scripts/timeline.py --plain 4545 4 > $O/${TAG}_full2.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_umma -s 26 -c 2 -o $O/${TAG}_full_attn_c3 -f python scripts/timeline.py --plain 4545 4 > $O/${TAG}_full3.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_flash -c 2 -o $O/${TAG}_full_attn_flash -f python scripts/attn_bench.py --once --shapes 4545,8,128,0 4545,8,64,0 > $O/${TAG}_full4.log 2>&1
python - <<PY
import json
for f in ("config3","config2","config5","strong_n1","config3_fusedtr","config2_fusedtr"):
    try:
        d=json.loads(open("$O/${TAG}_bench_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "ms/step %.4f" % d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("ms_per_step"), "frac", (d.get("roofline") or {}).get("frac"), "parity", (d.get("parity") or {}).get("rel_l2"), "clk", (d.get("clocks") or {}).get("samples"))
    except Exception as e: print(f, "ERR", e)
PY
cut -c1-200 $O/${TAG}_attn_bench.jsonl
