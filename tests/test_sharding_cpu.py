"""Host-side sharding logic, no GPU: shard ranges, RNG slicing, and a world_size-2 `gloo`
run of the sharded sampler against the unsharded run.  The denoiser in these tests is the CPU oracle (tests may use
oracle/ as the checker); the product model has no CPU path.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jen1_b200.config import tiny_desc
from jen1_b200.diffusion import create_gaussian_diffusion
from jen1_b200.sharding import gather_latents, shard_conditioning, shard_range, sharded_sample
from jen1_b200.weights import random_state_dict
from oracle.make_golden import make_inputs
from oracle.unet_oracle import OracleUNet


def test_shard_range_partitions_exactly():
    for n in (0, 1, 2, 3, 7, 8, 16, 32, 33):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans[:-1], spans[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sorted(sizes, reverse=True) == sizes
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_shard_conditioning_keeps_none_and_slices_batch():
    cond = dict(cross_attn_cond=torch.arange(12.).reshape(4, 3, 1), cross_attn_masks=torch.ones(4, 3, dtype=torch.bool),
                global_cond=None, input_concat_cond=torch.arange(8.).reshape(4, 2, 1))
    s = shard_conditioning(cond, 1, 3)
    assert s["global_cond"] is None and s["cross_attn_cond"].shape[0] == 2
    assert torch.equal(s["input_concat_cond"], cond["input_concat_cond"][1:3])


def test_sharded_rng_draws_are_slices_of_the_full_batch_draw():
    d = create_gaussian_diffusion(steps=1000, noise_schedule="linear", objective="noise", device="cpu",
                                  cfg_dropout_proba=0.5, embedding_scale=0.8, sampling_steps=20)
    torch.manual_seed(5)
    full_n, full_b = d._randn((6, 4, 9), "cpu"), d._bernoulli(6, "cpu")
    for lo, hi in ((0, 2), (2, 6), (5, 6)):
        d.shard = (lo, hi, 6)
        torch.manual_seed(5)
        n, b = d._randn((hi - lo, 4, 9), "cpu"), d._bernoulli(hi - lo, "cpu")
        assert torch.equal(n, full_n[lo:hi]) and torch.equal(b, full_b[lo:hi])
    d.shard = None


def _problem():
    desc = tiny_desc()
    sd = random_state_dict(desc, 7)
    B, T, S = 3, 20, 20
    x, t, emb, mask, cc = make_inputs(desc, B, T, 31, 2)
    cond = dict(cross_attn_cond=emb, cross_attn_masks=mask, global_cond=None, input_concat_cond=cc)
    # cfg_dropout_proba=0: on the generic (callable) loop the bernoulli is drawn inside the model, per shard
    d = create_gaussian_diffusion(steps=1000, noise_schedule="linear", objective="noise", device="cpu",
                                  cfg_dropout_proba=0.0, embedding_scale=0.8, batch_cfg=True, scale_cfg=True,
                                  sampling_steps=S)
    return desc, OracleUNet(desc, sd), d, cond, (B, desc.in_channels, T)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        desc, model, d, cond, shape = _problem()
        torch.manual_seed(123)
        local = sharded_sample(d, model, shape, cond, rank, world, init_data=torch.zeros(shape))
        lo, hi = shard_range(shape[0], rank, world)
        assert local.shape[0] == hi - lo
        full = gather_latents(local, shape[0])
        if rank == 0:
            q.put(full)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_world2_gloo_sharded_sample_equals_unsharded():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=500)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    desc, model, d, cond, shape = _problem()
    torch.manual_seed(123)
    ref = d.sample(model, shape, cond, init_data=torch.zeros(shape))
    assert got.shape == ref.shape
    # samples are independent: per-sample arithmetic is identical up to batched-GEMM blocking on the CPU
    assert (got - ref).abs().max().item() < 1e-4, (got - ref).abs().max().item()
