#!/bin/bash
# A/B of library variants (scripts/build_variant.py): op-by-op check + the two quick benches for each.
# Usage (under gpurun): bash scripts/gpu_variants.sh <tag> <variant|tree> ...
TAG=$1; shift
O=gpurun_out
mkdir -p $O
# a variant may carry one environment setting: name@VAR=value
for VV in "$@"; do
  V=${VV%%@*}; E=${VV#*@}; [ "$E" = "$VV" ] && E="JEN1_NOP=1"
  export $E
  if [ "$V" = tree ]; then unset JEN1_B200_LIB; else export JEN1_B200_LIB=$PWD/jen1_b200/_C/variants/$V/libjen1_b200.so; fi
  if [ "$V" = tree ] || [ -n "$AB" ]; then
    ( timeout 200 python scripts/umma_debug.py 150 2 cfg ) > $O/${TAG}_${VV}_ab.log 2>&1 || echo "$V: A/B FAILED"
    tail -1 $O/${TAG}_${VV}_ab.log
  fi
  for W in config3 config2; do
    timeout 300 python bench.py --workload $W --steps 60 --warmup 5 --quick --no-cpu-baseline --no-e2e > $O/${TAG}_${VV}_$W.json 2> $O/${TAG}_${VV}_$W.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_${VV}_$W.json").read().strip().splitlines()[-1]); print("$VV $W ms/step %.4f" % d["ms_per_step"])
except Exception as e: print("$V $W ERR", e)
PY
  done
  unset ${E%%=*}
done
if [ -n "$TL" ]; then
  unset JEN1_B200_LIB
  JEN1_TIMELINE=1 JEN1_TRACE=1 timeout 120 python scripts/timeline.py 1515 1 > /dev/null 2> $O/${TAG}_timeline_c2.raw
  JEN1_TIMELINE=1 JEN1_TRACE=1 timeout 120 python scripts/timeline.py 4545 4 > /dev/null 2> $O/${TAG}_timeline_c3.raw
  python scripts/tl_table.py $O/${TAG}_timeline_c2.raw > $O/${TAG}_timeline_c2.txt 2>&1
  python scripts/tl_table.py $O/${TAG}_timeline_c3.raw > $O/${TAG}_timeline_c3.txt 2>&1
  tail -12 $O/${TAG}_timeline_c3.txt
fi
