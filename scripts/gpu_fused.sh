#!/bin/bash
# A/B of the fused Transformer1d kernel (JEN1_FUSED_TR=1) against the unfused chain.  Usage (under gpurun): bash scripts/gpu_fused.sh <tag>
TAG=${1:-fz}
O=gpurun_out
mkdir -p $O
export JEN1_FUSED_TR=1
( timeout 200 python scripts/umma_debug.py 150 2 cfg ) > $O/${TAG}_ab_full.log 2>&1 || { echo "A/B full FAILED/HUNG"; tail -12 $O/${TAG}_ab_full.log; exit 1; }
tail -1 $O/${TAG}_ab_full.log
( timeout 200 python scripts/umma_debug.py 333 1 causal ) > $O/${TAG}_ab_causal.log 2>&1; tail -1 $O/${TAG}_ab_causal.log
( timeout 300 python scripts/umma_debug.py 4545 4 cfg ) > $O/${TAG}_ab_c3.log 2>&1; tail -1 $O/${TAG}_ab_c3.log
timeout 300 python -m pytest tests/test_umma_gpu.py tests/test_bench_configs_gpu.py -m gpu -x -q --timeout 300 -k "not trajectory" > $O/${TAG}_pytest.log 2>&1; tail -3 $O/${TAG}_pytest.log
timeout 300 python bench.py --steps 50 --warmup 5 --quick --no-cpu-baseline --no-e2e > $O/${TAG}_bench_config3.json 2> $O/${TAG}_bench_config3.err
timeout 300 python bench.py --workload config2 --steps 50 --warmup 5 --quick --no-cpu-baseline --no-e2e > $O/${TAG}_bench_config2.json 2> $O/${TAG}_bench_config2.err
JEN1_TIMELINE=1 timeout 120 python scripts/timeline.py 1515 1 > /dev/null 2> $O/${TAG}_timeline_c2.raw
JEN1_TIMELINE=1 timeout 120 python scripts/timeline.py 4545 4 > /dev/null 2> $O/${TAG}_timeline_c3.raw
python - <<PY
import json
for f in ("$O/${TAG}_bench_config3.json","$O/${TAG}_bench_config2.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "ms/step", d["ms_per_step"], "launches", d["launches_per_step"])
    except Exception as e: print(f, "ERR", e)
for f in ("$O/${TAG}_timeline_c2.raw","$O/${TAG}_timeline_c3.raw"):
    txt=open(f).read().split('==== second (warm) evaluation')[1]
    lines=[l for l in txt.splitlines() if 'trtl' in l]
    print(f, len(lines))
    for l in lines[20:30]+lines[100:110]: print(l[11:])
PY
