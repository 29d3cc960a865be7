#!/bin/bash
# N-GPU pass (gpurun --gpus N): NCCL gather check, weak and strong scaling bench lines.  Usage: bash scripts/gpu_multi.sh <tag> <N>
TAG=${1:-mg}; N=${2:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 scripts/multigpu_check.py > $O/${TAG}_multigpu_check.log 2>&1; echo "check rc=$?"; grep multigpu_check $O/${TAG}_multigpu_check.log
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 50 --warmup 5 > $O/${TAG}_bench_config3_n$N.json 2> $O/${TAG}_bench_n$N.err; echo "weak rc=$?"
timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 50 --warmup 5 --scaling strong > $O/${TAG}_bench_strong_n$N.json 2> $O/${TAG}_bench_strong_n$N.err; echo "strong rc=$?"
timeout 600 $TR --master-port 29514 bench.py --gpus $N --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_n$N.json 2> /dev/null; echo "ref rc=$?"
python - <<PY
import json
for f in ("config3_n$N","strong_n$N","reference_n$N"):
    try:
        d=json.loads(open("$O/${TAG}_bench_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "n_gpus", d["n_gpus"], "ms/step %.4f" % d["ms_per_step"], "value %.0f" % d["value"], "scaling", d["scaling"], "e2e", (d.get("e2e") or {}).get("ms_per_step"))
    except Exception as e: print(f, "ERR", e)
PY
