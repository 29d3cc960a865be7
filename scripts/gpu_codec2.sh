#!/bin/bash
TAG=${1:-k}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_codec_gpu.py tests/test_generation_gpu.py -m gpu -x -q --timeout 300 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 600 python bench.py --steps 30 --warmup 3 --no-gpu-eager > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -2 $O/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1]); print(d["ms_per_step"], json.dumps(d["extra"]["codec_decode"], indent=1))
PY
