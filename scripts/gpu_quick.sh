#!/bin/bash
# Quick GPU pass while iterating on kernels: A/B op-by-op (with and without the tcgen05 attention), GPU tests,
# a short bench and a timeline.  Every stage has a tight timeout; a failed A/B stops the pass.
# Usage (under gpurun): bash scripts/gpu_quick.sh <tag> [skip_convonly]
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
if [ -z "$2" ]; then
( JEN1_ATTN_IMPL=fma timeout 150 python scripts/umma_debug.py 150 2 cfg ) > $O/${TAG}_ab_convonly.log 2>&1 || { echo "A/B conv-only FAILED/HUNG"; tail -5 $O/${TAG}_ab_convonly.log; exit 1; }
tail -2 $O/${TAG}_ab_convonly.log
fi
( timeout 150 python scripts/umma_debug.py 150 2 cfg ) > $O/${TAG}_ab_full.log 2>&1 || { echo "A/B full FAILED/HUNG"; tail -12 $O/${TAG}_ab_full.log; exit 1; }
tail -8 $O/${TAG}_ab_full.log
( timeout 150 python scripts/umma_debug.py 333 1 causal ) > $O/${TAG}_ab_causal.log 2>&1; tail -3 $O/${TAG}_ab_causal.log
timeout 500 python -m pytest tests -m gpu -x -q --timeout 200 > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
tail -5 $O/${TAG}_pytest_gpu.log
timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_config2.json 2> $O/${TAG}_bench_config2.err
timeout 200 python bench.py --workload config3 --steps 20 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_config3.json 2> $O/${TAG}_bench_config3.err
JEN1_TIMELINE=1 timeout 120 python scripts/timeline.py 1515 1 > /dev/null 2> $O/${TAG}_timeline_c2.txt
JEN1_TIMELINE=1 timeout 120 python scripts/timeline.py 4545 4 > /dev/null 2> $O/${TAG}_timeline_c3.txt
python - <<PY
import json
for f in ("$O/${TAG}_bench_config2.json","$O/${TAG}_bench_config3.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 $O/${TAG}_bench_config2.err
grep "total:" $O/${TAG}_timeline_c2.txt $O/${TAG}_timeline_c3.txt
