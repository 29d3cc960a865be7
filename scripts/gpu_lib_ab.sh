#!/bin/bash
# Library-variant A/B inside one GPU call: quick benches of both workloads for the in-tree library (first and last) and
# every variant built with scripts/build_variant.py.  Usage (under gpurun): bash scripts/gpu_lib_ab.sh <tag> <variant> ...
TAG=${1:-lv}; shift
O=gpurun_out
mkdir -p $O
run() {
  local name=$1 lib=$2
  for wl in config3 config2; do
    JEN1_B200_LIB=$lib timeout 200 python bench.py --workload $wl --steps 40 --warmup 5 --quick --no-cpu-baseline --no-e2e > $O/${TAG}_${name}_$wl.json 2> /dev/null
    python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_${name}_$wl.json").read().strip().splitlines()[-1]); print("$name $wl ms/step %.4f parity %s" % (d["ms_per_step"], (d.get("parity") or {}).get("rel_l2")))
except Exception as e: print("$name $wl ERR", e)
PY
  done
}
run base jen1_b200/_C/libjen1_b200.so
for v in "$@"; do run $v jen1_b200/_C/variants/$v/libjen1_b200.so; done
run base2 jen1_b200/_C/libjen1_b200.so
