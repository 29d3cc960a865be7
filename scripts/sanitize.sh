#!/bin/bash
# compute-sanitizer pass over the hand-written kernels + split-K determinism stress.
# Usage (under gpurun): bash scripts/sanitize.sh <tag>     -> gpurun_out/<tag>_sanitize_*.log
TAG=${1:-san}
O=gpurun_out
mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 $CS --tool $tool --print-limit 20 python scripts/sanitize_driver.py tiny > $O/${TAG}_sanitize_${tool}_tiny.log 2>&1
  echo "$tool tiny rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/${TAG}_sanitize_${tool}_tiny.log | tail -1)"
done
for tool in memcheck racecheck synccheck; do
  timeout 900 $CS --tool $tool --print-limit 20 python scripts/sanitize_driver.py codec > $O/${TAG}_sanitize_${tool}_codec.log 2>&1
  echo "$tool codec rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/${TAG}_sanitize_${tool}_codec.log | tail -1)"
done
timeout 900 $CS --tool memcheck --print-limit 20 python scripts/sanitize_driver.py full > $O/${TAG}_sanitize_memcheck_full.log 2>&1
echo "memcheck full rc=$? : $(grep -E 'ERROR SUMMARY' $O/${TAG}_sanitize_memcheck_full.log | tail -1)"
timeout 900 $CS --tool racecheck --print-limit 20 python scripts/sanitize_driver.py full > $O/${TAG}_sanitize_racecheck_full.log 2>&1
echo "racecheck full rc=$? : $(grep -E 'RACECHECK SUMMARY' $O/${TAG}_sanitize_racecheck_full.log | tail -1)"
timeout 600 python scripts/sanitize_driver.py stress 2000 > $O/${TAG}_stress.log 2>&1; tail -1 $O/${TAG}_stress.log
