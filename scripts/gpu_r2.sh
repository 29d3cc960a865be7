#!/bin/bash
# Round-2 GPU pass: GPU tests, the default bench (config 3) + reference arm, timelines.  Usage (under gpurun): bash scripts/gpu_r2.sh <tag> [notests]
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
if [ -z "$2" ]; then
timeout 1200 python -m pytest tests -m gpu -x -q -s --timeout 600 > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
grep -E "rel-L2|passed|failed|rc=|Error|error" $O/${TAG}_pytest_gpu.log | tail -30
fi
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_config3.json 2> $O/${TAG}_bench_config3.err; echo "bench rc=$?"
timeout 300 python bench.py --workload config2 --steps 50 --warmup 5 --quick --no-cpu-baseline > $O/${TAG}_bench_config2.json 2> $O/${TAG}_bench_config2.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $O/${TAG}_bench_reference_config3.json 2> $O/${TAG}_bench_reference.err
JEN1_TIMELINE=1 timeout 120 python scripts/timeline.py 1515 1 > /dev/null 2> $O/${TAG}_timeline_c2.raw
JEN1_TIMELINE=1 timeout 120 python scripts/timeline.py 4545 4 > /dev/null 2> $O/${TAG}_timeline_c3.raw
python scripts/tl_table.py $O/${TAG}_timeline_c2.raw > $O/${TAG}_timeline_c2.txt 2>&1
python scripts/tl_table.py $O/${TAG}_timeline_c3.raw > $O/${TAG}_timeline_c3.txt 2>&1
python - <<PY
import json
for f in ("$O/${TAG}_bench_config3.json","$O/${TAG}_bench_config2.json","$O/${TAG}_bench_reference_config3.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("ms_per_step"), "parity", d.get("parity"), "eager", d.get("gpu_eager_baseline"), "extra", d.get("extra"), "clocks", d.get("clocks"))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 $O/${TAG}_bench_config3.err
tail -20 $O/${TAG}_timeline_c3.txt
