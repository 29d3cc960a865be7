#!/bin/bash
# gpu_ab.sh + the fused-transformer arm + the attention tests.  Usage (under gpurun): bash scripts/gpu_ab2.sh <tag>
TAG=${1:-ab}
O=gpurun_out
bash scripts/gpu_ab.sh $TAG
timeout 300 python -m pytest tests/test_attention_gpu.py tests/test_engine_gpu.py -m gpu -x -q --timeout 300 > $O/${TAG}_pytest2.log 2>&1; tail -2 $O/${TAG}_pytest2.log
JEN1_FUSED_TR=1 timeout 300 python bench.py --steps 50 --warmup 5 --quick --no-cpu-baseline --no-e2e > $O/${TAG}_bench_config3_fusedtr.json 2> /dev/null
JEN1_FUSED_TR=1 timeout 300 python bench.py --workload config2 --steps 50 --warmup 5 --quick --no-cpu-baseline --no-e2e > $O/${TAG}_bench_config2_fusedtr.json 2> /dev/null
python - <<PY
import json
for f in ("config3_fusedtr","config2_fusedtr"):
    try:
        d=json.loads(open("$O/${TAG}_bench_%s.json" % f).read().strip().splitlines()[-1]); print(f, "ms/step", d["ms_per_step"], "parity", (d.get("parity") or {}).get("rel_l2"))
    except Exception as e: print(f, "ERR", e)
PY
