#!/bin/bash
# Re-entry sanity pass: GPU tests, smoke, the default bench line.  Usage (under gpurun): bash scripts/gpu_sanity.sh <tag>
TAG=${1:-san}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
tail -4 $O/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -3 $O/${TAG}_smoke.log
timeout 600 python bench.py > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err; echo "bench rc=$?"
cat $O/${TAG}_bench_default.json | cut -c1-1500
