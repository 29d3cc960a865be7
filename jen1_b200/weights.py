"""Weights for the denoiser: deterministic random initialisation and reference-checkpoint ingestion.

The reference ships no checkpoint, so tests and the benchmark use seeded random-init
weights of the reference architecture.  `random_state_dict` draws every tensor of `UNetDesc.tensor_spec()`
from one CPU generator in spec order, with PyTorch-default-like scales (U(-1/sqrt(fan_in), 1/sqrt(fan_in))
for conv/linear weights and biases, N(0,1) for the embedding table and the learned Fourier frequencies) and
non-trivial norm affines so that gamma/beta handling is exercised.  The same dict loads into the reference
`UNetCFG1d` via `load_state_dict` (used by oracle/make_golden.py) and into the B200 engine.

`load_checkpoint_state_dict` reads the reference's checkpoint file layout
({'model','epoch','optimizer','learning_rate'}, optional `_orig_mod.` prefixes -- reference
utils/script_util.py:85-88, 93-122).
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from .config import UNetDesc


def random_state_dict(desc: UNetDesc, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    sd: Dict[str, torch.Tensor] = {}
    for name, shape, kind in desc.tensor_spec():
        if kind.startswith(("conv_w:", "linear_w:", "bias:")):
            fan_in = int(kind.split(":")[1])
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound
        elif kind == "norm_w":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
        elif kind == "norm_b":
            t = 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
        elif kind in ("embedding", "posemb"):
            t = torch.randn(shape, generator=g, dtype=torch.float32)
        else:  # pragma: no cover
            raise ValueError(kind)
        sd[name] = t.to(dtype)
    return sd


def check_state_dict(desc: UNetDesc, sd: Dict[str, torch.Tensor]) -> None:
    """Raise if `sd` is not a complete UNetCFG1d state_dict for `desc`."""
    spec = desc.tensor_spec()
    missing = [n for n, _, _ in spec if n not in sd]
    if missing:
        raise KeyError("state_dict is missing %d tensors, first: %s" % (len(missing), missing[:3]))
    for n, shape, _ in spec:
        if tuple(sd[n].shape) != tuple(shape):
            raise ValueError("tensor %s has shape %s, expected %s" % (n, tuple(sd[n].shape), tuple(shape)))


def load_checkpoint_state_dict(path: str, desc: UNetDesc) -> Dict[str, torch.Tensor]:
    """Read a reference checkpoint and return a clean UNetCFG1d state_dict (CPU fp32)."""
    blob = torch.load(path, map_location="cpu")
    saved = blob["model"] if isinstance(blob, dict) and "model" in blob else blob
    out = {}
    for name, _, _ in desc.tensor_spec():
        if name in saved:
            out[name] = saved[name]
        elif "_orig_mod." + name in saved:
            out[name] = saved["_orig_mod." + name]
        elif "module." + name in saved:
            out[name] = saved["module." + name]
    check_state_dict(desc, out)
    return {k: v.detach().to(torch.float32).contiguous() for k, v in out.items()}
