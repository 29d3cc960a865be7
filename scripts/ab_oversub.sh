# A/B sweep of the oversubscription penalties of the tile planner (same box, back to back)
for cfg in "1.15 1.5" "1.0 1.0" "1.0 0.3" "0.85 0.0"; do
  set -- $cfg
  for wl in config3 config2; do
    JEN1_OVERSUB_MUL=$1 JEN1_OVERSUB_ADD=$2 timeout 200 python bench.py --workload $wl --steps 30 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r01u_${wl}_$1_$2.json 2> gpurun_out/r01u_${wl}_$1_$2.err
    python -c "
import json,sys
try: print('$wl mul=$1 add=$2', json.loads(open('gpurun_out/r01u_${wl}_$1_$2.json').read().strip().splitlines()[-1])['ms_per_step'])
except Exception as e: print('$wl $1 $2 ERR', e)"
  done
done
