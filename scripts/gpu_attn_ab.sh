#!/bin/bash
# Attention-kernel iteration pass: parity tests of the attention kernels, then scripts/attn_bench.py for the in-tree
# library and every variant library given.  Usage (under gpurun): bash scripts/gpu_attn_ab.sh <tag> [variant ...]
TAG=${1:-fa}; shift
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_attention_gpu.py -m gpu -x -q --timeout 120 > $O/${TAG}_pytest_attn.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_attn.log
tail -4 $O/${TAG}_pytest_attn.log
SH="4545,8,128,0 4545,8,128,1 4545,8,64,0 4545,8,64,1 1137,8,64,0"
echo "== base"; timeout 120 python scripts/attn_bench.py --shapes $SH 2> $O/${TAG}_base.err | tee $O/${TAG}_attn_base.jsonl | cut -c1-175
for v in "$@"; do
  echo "== $v"
  JEN1_B200_LIB=jen1_b200/_C/variants/$v/libjen1_b200.so timeout 120 python scripts/attn_bench.py --shapes $SH 2> $O/${TAG}_$v.err | tee $O/${TAG}_attn_$v.jsonl | cut -c1-175
done
