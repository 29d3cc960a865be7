"""Shard-vs-whole equality on the GPU ("multi-GPU without a cluster"): sampling a batch of 4 on
one device must equal sampling two shards of 2 with the full-batch random draws sliced per shard
(jen1_b200/sharding.py).  Samples are independent; the only difference is how tiles / split-K fold the batch rows,
i.e. floating-point summation order, so the comparison is tight in fp32 strict mode.
"""
import pytest
import torch

from jen1_b200.config import tiny_desc
from jen1_b200.diffusion import create_gaussian_diffusion
from jen1_b200.sharding import shard_range, sharded_sample
from jen1_b200.weights import random_state_dict
from oracle.make_golden import make_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("dtype,tol", [("fp32", 2e-3), ("bf16", 6e-2)])
def test_sharded_sampling_matches_unsharded(dtype, tol):
    from jen1_b200.model import UNetCFG1d
    desc = tiny_desc()
    sd = random_state_dict(desc, 7)
    model = UNetCFG1d(desc, device=DEV, dtype=dtype).load_state_dict(sd)
    B, T, S = 4, 50, 20
    _, _, emb, mask, cc = make_inputs(desc, B, T, 41, 3)
    cond = dict(cross_attn_cond=emb.to(DEV), cross_attn_masks=mask.to(DEV), global_cond=None,
                input_concat_cond=cc.to(DEV))
    d = create_gaussian_diffusion(steps=1000, noise_schedule="linear", objective="noise", device=DEV,
                                  cfg_dropout_proba=0.2, embedding_scale=0.8, batch_cfg=True, scale_cfg=True,
                                  sampling_steps=S, rng_device="cpu")
    shape = (B, desc.in_channels, T)
    torch.manual_seed(5)
    whole = d.sample(model, shape, cond).cpu()
    parts = []
    for rank in range(2):
        torch.manual_seed(5)
        part = sharded_sample(d, model, shape, cond, rank, 2).cpu()
        lo, hi = shard_range(B, rank, 2)
        assert part.shape[0] == hi - lo
        parts.append(part)
    got = torch.cat(parts, 0)
    err = ((got - whole).norm() / whole.norm()).item()
    assert err < tol, err


def test_nccl_gather_of_sharded_latents_across_gpus():
    """The optional NCCL/NVLink gather of finished latents (jen1_b200/sharding.py:gather_latents) on real GPUs: torchrun
    with one rank per GPU, uneven shards, gathered result == unsharded run (scripts/multigpu_check.py).  Needs >= 2 GPUs."""
    import os
    import subprocess
    import sys
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 4)), "--master-addr",
           "127.0.0.1", "--master-port", "29531", os.path.join(root, "scripts", "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "rel-L2 vs unsharded run" in r.stdout
