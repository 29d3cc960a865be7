#!/bin/bash
# ncu full-set captures of chosen conv_umma launches of one UNet evaluation (second, warm evaluation of scripts/timeline.py).
# Usage: bash scripts/gpu_ncu.sh <tag> <T> <B> <op index> [count]
TAG=$1; T=$2; B=$3; OP=$4; CNT=${5:-2}
O=gpurun_out
mkdir -p $O
SKIP=$((234 + OP))
JEN1_TIMELINE= timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s $SKIP -c $CNT -o $O/${TAG} -f python scripts/timeline.py --plain $T $B > $O/${TAG}.log 2>&1
tail -3 $O/${TAG}.log
