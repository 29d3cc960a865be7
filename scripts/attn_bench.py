#!/usr/bin/env python
"""Tensor-pipe evidence for the attention kernels on stated synthetic shapes (SURVEY.md section 7 hard part 2): times
`jen1_attention_forward` with CUDA events and prints TFLOP/s (4*B*H*N*M*d, halved when causal) against the measured bf16
peak.  `--once` runs a single launch per shape (for `ncu --set full -k regex:attn_flash`).

    python scripts/attn_bench.py [--once] [--shapes N,H,d,causal ...]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jen1_b200.config import tiny_desc  # noqa: E402
from jen1_b200.model import UNetCFG1d  # noqa: E402
from jen1_b200.weights import random_state_dict  # noqa: E402

DEFAULT = ["4545,8,64,0", "4545,8,64,1", "4545,8,128,0", "4545,8,128,1", "1137,8,64,0", "285,8,64,0", "72,8,32,0"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--shapes", nargs="*", default=DEFAULT)
    args = ap.parse_args()
    dev = "cuda:0"
    desc = tiny_desc()
    eng = UNetCFG1d(desc, device=dev, dtype="bf16").load_state_dict(random_state_dict(desc, 7)).engine
    peak = 1590.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk)).get("bf16_tflops", peak))
    B = args.batch
    for spec in args.shapes:
        N, H, d, causal = [int(v) for v in spec.split(",")]
        qkv = torch.randn(B, N, 3 * H * d, device=dev).to(torch.bfloat16)
        impl = "tcgen05"
        eng.attention(qkv, H, causal=bool(causal), impl=impl)
        torch.cuda.synchronize()
        if args.once:
            continue
        for _ in range(3):
            eng.attention(qkv, H, causal=bool(causal), impl=impl)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            eng.attention(qkv, H, causal=bool(causal), impl=impl)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        flops = 4.0 * B * H * N * N * d * (0.5 if causal else 1.0)
        tf = flops / (ms * 1e-3) / 1e12
        print(json.dumps({"shape": {"B": B, "N": N, "M": N, "H": H, "d": d, "causal": bool(causal)},
                          "kernel": "attn_flash_kernel" if N > 256 else "attn_umma_kernel", "ms": ms, "tflops": tf,
                          "frac_of_measured_bf16_peak": tf / peak, "peak_tflops": peak}), flush=True)


if __name__ == "__main__":
    main()
