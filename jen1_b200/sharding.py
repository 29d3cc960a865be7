"""Batch sharding of the sampling hot path over the GPUs of one box.

The independent units are diffusion samples: rank r of W owns a contiguous slice of the global batch of
`x / cross_attn_cond / cross_attn_masks / input_concat_cond`, weights are replicated, and there is NO collective
inside or between sampler steps.  The only communication is an optional `all_gather` of the finished latents.

Seed parity with the single-device run: every rank draws the FULL-batch random tensors (initial `randn(shape)`,
per step the cond-dropout bernoulli `(B,1,1)` of reference jen1/model/model.py:325 and `randn_like` of
jen1/diffusion/gdm/gdm.py:218) from the same generator state and keeps its slice, so a W-way run reproduces the
latents of the unsharded run with the same seed (`GaussianDiffusion.shard`).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of `n` samples: the first `n % world` ranks own one extra sample."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world: %d/%d" % (rank, world))
    if n < 0:
        raise ValueError("negative batch")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_conditioning(conditioning: Dict[str, Optional[torch.Tensor]], lo: int, hi: int) -> Dict:
    """Slice the batch dimension of the reference conditioning dict (reference generation.py:187-192)."""
    out = {}
    for k, v in conditioning.items():
        out[k] = v[lo:hi] if torch.is_tensor(v) else v
    return out


def sharded_sample(diffusion, model, global_shape, conditioning, rank: int, world: int, *, causal: bool = False,
                   init_data: Optional[torch.Tensor] = None, return_all_timesteps: bool = False) -> torch.Tensor:
    """`diffusion.sample` for this rank's slice of a global batch; returns the LOCAL latents [hi-lo, C, T].

    `conditioning` and `init_data` are the GLOBAL tensors (every rank holds the same prompts; only its slice is
    moved through the model).  An empty shard (more ranks than samples) returns an empty tensor without touching
    the model, after consuming nothing from the RNG stream that other ranks depend on.
    """
    B = int(global_shape[0])
    lo, hi = shard_range(B, rank, world)
    if hi == lo:
        return torch.empty((0,) + tuple(global_shape[1:]), device=diffusion.device)
    prev = getattr(diffusion, "shard", None)
    diffusion.shard = (lo, hi, B)
    try:
        return diffusion.sample(model, (hi - lo,) + tuple(global_shape[1:]), shard_conditioning(conditioning, lo, hi),
                                return_all_timesteps=return_all_timesteps, causal=causal,
                                init_data=None if init_data is None else init_data[lo:hi])
    finally:
        diffusion.shard = prev


def gather_latents(local: torch.Tensor, global_batch: int, group=None) -> torch.Tensor:
    """Collect every rank's latents on every rank (`all_gather`; NCCL over NVLink on GPUs, gloo on CPU).

    Shards may be uneven: each rank contributes a buffer padded to the largest shard.  ~2.3 MB per 30 s sample.
    """
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        assert local.shape[0] == global_batch
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    most = shard_range(global_batch, 0, world)
    most = most[1] - most[0]
    pad = torch.zeros((most,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_range(global_batch, r, world)
        parts.append(bufs[r][: hi - lo])
    return torch.cat(parts, dim=0)
