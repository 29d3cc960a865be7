#!/bin/bash
# Planner A/B inside one GPU call: quick benches of both workloads for a list of "VAR=value" settings.
# Usage (under gpurun): bash scripts/gpu_plan_ab.sh <tag> "JEN1_KSTEP_US=0.45" "JEN1_KSTEP_US=0.3 JEN1_XCH_US=0.15" ...
TAG=${1:-pl}; shift
O=gpurun_out
mkdir -p $O
run() {
  local name=$1; shift
  for wl in config3 config2; do
    env "$@" timeout 200 python bench.py --workload $wl --steps 40 --warmup 5 --quick --no-cpu-baseline --no-e2e > $O/${TAG}_${name}_$wl.json 2> /dev/null
    python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_${name}_$wl.json").read().strip().splitlines()[-1]); print("$name $wl ms/step %.4f" % d["ms_per_step"])
except Exception as e: print("$name $wl ERR", e)
PY
  done
}
run base JEN1_NOOP=1
i=0
for setting in "$@"; do i=$((i+1)); echo "== v$i: $setting"; run v$i $setting; done
run base2 JEN1_NOOP=1
