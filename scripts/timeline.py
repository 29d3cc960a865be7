#!/usr/bin/env python
"""Print per-op phase clocks of the tcgen05 conv kernel for one UNet evaluation (JEN1_TIMELINE debugging aid).

    python scripts/timeline.py [--plain] [T = 1515] [B = 1]
"""
import os
import sys
PLAIN = "--plain" in sys.argv  # no in-kernel clocks / plan trace: the bare two-evaluation workload for ncu captures
if PLAIN:
    sys.argv.remove("--plain")
else:
    os.environ["JEN1_TIMELINE"] = "1"
    os.environ["JEN1_TRACE"] = "1"
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jen1_b200.config import UNetDesc
from jen1_b200.model import UNetCFG1d
from jen1_b200.weights import random_state_dict

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1515
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
desc = UNetDesc()
m = UNetCFG1d(desc, device="cuda:0", dtype="bf16").load_state_dict(random_state_dict(desc, 0))
g = torch.Generator().manual_seed(5)
x = torch.randn(B, 128, T, generator=g).cuda()
t = torch.randint(0, 1000, (B,), generator=g).cuda()
emb = torch.randn(B, 128, 1024, generator=g).cuda()
mask = torch.ones(B, 128, dtype=torch.bool).cuda()
cc = torch.randn(B, 129, T, generator=g).cuda()
for it in range(2):
    if it == 1:
        print("==== second (warm) evaluation", file=sys.stderr, flush=True)
    y = m(x, t, embedding=emb, embedding_mask=mask, features=None, channels_list=[cc], embedding_scale=0.8,
          batch_cfg=True, scale_cfg=True)
    torch.cuda.synchronize()
