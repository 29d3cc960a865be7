#!/usr/bin/env python
"""A/B the tcgen05 conv path against the generic fp32-FMA kernel, op by op, on the full model (GPU only).

    python scripts/umma_debug.py [T] [B] [variant]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jen1_b200.config import UNetDesc  # noqa: E402
from jen1_b200.model import UNetCFG1d  # noqa: E402
from jen1_b200.weights import random_state_dict  # noqa: E402


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def make(desc, sd, impl):
    os.environ["JEN1_CONV_IMPL"] = impl
    m = UNetCFG1d(desc, device="cuda:0", dtype="bf16").load_state_dict(sd)
    os.environ.pop("JEN1_CONV_IMPL")
    return m


def compare(T=150, B=2, variant="cfg", verbose=True, models=None, masked_tail=0):
    desc = UNetDesc()
    if models is None:
        sd = random_state_dict(desc, 0)
        models = (make(desc, sd, "generic"), make(desc, sd, "umma"))
    mg, mu = models
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, 128, T, generator=g).cuda()
    t = torch.randint(0, 1000, (B,), generator=g).cuda()
    emb = torch.randn(B, 128, 1024, generator=g).cuda()
    mask = torch.ones(B, 128, dtype=torch.bool)
    if masked_tail:  # what T5Conditioner produces for short prompts: padded rows masked out and zeroed
        mask[:, 128 - masked_tail:] = False
        emb = emb * mask[:, :, None].cuda()
    mask = mask.cuda()
    cc = torch.randn(B, 129, T, generator=g).cuda()
    kw = {"plain": dict(embedding_scale=1.0),
          "cfg": dict(embedding_scale=0.8, batch_cfg=True, scale_cfg=True),
          "causal": dict(embedding_scale=0.8, batch_cfg=True, scale_cfg=True, causal=True)}[variant]
    outs = []
    for m in (mg, mu):
        y = m(x, t, embedding=emb, embedding_mask=mask, features=None, channels_list=[cc], **kw)
        torch.cuda.synchronize()
        outs.append(y.cpu())
    worst, nbad, i = 0.0, 0, 0
    for pref in ("at", "op"):
      i = 0
      while True:
        name = "%s%03d" % (pref, i)
        try:
            a = mg.engine.debug_tensor(name)
            b = mu.engine.debug_tensor(name)
        except Exception:
            break
        e = rel(b, a) if a.shape == b.shape else float("inf")
        fin = bool(torch.isfinite(b).all())
        worst = max(worst, e if fin else float("inf"))
        bad = (not fin) or e > 2e-2
        nbad += bad
        if verbose and (bad or i < 4):
            print("%s shape %s rel-L2 umma vs generic %.3e finite=%s%s" % (name, tuple(a.shape), e, fin, "  <-- BAD" if bad else ""))
        i += 1
    ef = rel(outs[1], outs[0])
    print("T=%d B=%d %s: %d ops compared, %d bad, worst %.3e, final output rel-L2 %.3e" % (T, B, variant, i, nbad, worst, ef))
    return nbad, worst, ef, models


if __name__ == "__main__":
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    v = sys.argv[3] if len(sys.argv) > 3 else "cfg"
    compare(T, B, v)
