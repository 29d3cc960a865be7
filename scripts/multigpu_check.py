#!/usr/bin/env python
"""Multi-GPU check under torchrun (one rank per GPU, NCCL): each rank samples its shard of a global batch with the
full-batch random draws sliced per shard, `gather_latents` (NCCL all_gather over NVLink) collects them, and rank 0
compares the gathered result with the unsharded run of the same seed on its own GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jen1_b200.config import tiny_desc  # noqa: E402
from jen1_b200.diffusion import create_gaussian_diffusion  # noqa: E402
from jen1_b200.model import UNetCFG1d  # noqa: E402
from jen1_b200.sharding import gather_latents, shard_range, sharded_sample  # noqa: E402
from jen1_b200.weights import random_state_dict  # noqa: E402
from oracle.make_golden import make_inputs  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    desc = tiny_desc()
    sd = random_state_dict(desc, 7)
    model = UNetCFG1d(desc, device=dev, dtype="fp32").load_state_dict(sd)
    B, T, S = 2 * world + 1, 50, 10  # uneven shards on purpose
    _, _, emb, mask, cc = make_inputs(desc, B, T, 41, 3)
    cond = dict(cross_attn_cond=emb.to(dev), cross_attn_masks=mask.to(dev), global_cond=None, input_concat_cond=cc.to(dev))
    d = create_gaussian_diffusion(steps=1000, noise_schedule="linear", objective="noise", device=dev, cfg_dropout_proba=0.2,
                                  embedding_scale=0.8, batch_cfg=True, scale_cfg=True, sampling_steps=S, rng_device="cpu")
    shape = (B, desc.in_channels, T)
    torch.manual_seed(5)
    part = sharded_sample(d, model, shape, cond, rank, world)
    lo, hi = shard_range(B, rank, world)
    assert part.shape[0] == hi - lo
    full = gather_latents(part, B)
    assert full.shape == shape
    ok = torch.ones(1, device=dev)
    if rank == 0:
        torch.manual_seed(5)
        whole = d.sample(model, shape, cond)
        err = ((full - whole).norm() / whole.norm()).item()
        print("multigpu_check: world %d, global batch %d (uneven shards), NCCL all_gather of latents, rel-L2 vs unsharded run %.3e"
              % (world, B, err), flush=True)
        ok[0] = 1.0 if err < 2e-3 else 0.0
    dist.broadcast(ok, 0)
    dist.destroy_process_group()
    sys.exit(0 if ok.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
