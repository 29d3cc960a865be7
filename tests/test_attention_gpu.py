"""GPU parity of the attention kernels on their own (C ABI `jen1_attention_forward`): the single-tile tcgen05 kernel
(<= 256 keys, the model's shapes), the key-tiled online-softmax tcgen05 kernel (any length; the synthetic long-sequence
shapes of the tensor-pipe evidence) and the fp32-FMA core, against a torch fp32 reference of reference
jen1/model/blocks.py:355-380 (softmax(q k^T d^-1/2 [+ causal mask]) v per head) on the SAME bf16 inputs.

Tolerance: the kernels round the softmax numerators P to bf16 before the P V product (fp32 accumulate) and the output to
bf16: rel-L2 <= 1e-2 (measured ~3e-3).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def engine():
    from jen1_b200.config import tiny_desc
    from jen1_b200.model import UNetCFG1d
    from jen1_b200.weights import random_state_dict
    desc = tiny_desc()
    return UNetCFG1d(desc, device=DEV, dtype="bf16").load_state_dict(random_state_dict(desc, 7)).engine


def torch_reference(qkv, heads, causal):
    B, N, C3 = qkv.shape
    C = C3 // 3
    d = C // heads
    q, k, v = [t.float().view(B, N, heads, d).transpose(1, 2) for t in qkv.split(C, dim=-1)]
    sim = torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5)
    if causal:
        keep = ~torch.ones((N, N), dtype=torch.bool, device=qkv.device).triu(1)
        sim = sim.masked_fill(~keep, -torch.finfo(sim.dtype).max)
    out = torch.matmul(sim.softmax(dim=-1, dtype=torch.float32), v)
    return out.transpose(1, 2).reshape(B, N, C)


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("B,N,H,d,causal,impl", [
    (2, 72, 8, 32, False, "tcgen05"), (2, 72, 8, 32, True, "tcgen05"), (2, 72, 8, 32, True, "flash"),
    (3, 129, 4, 16, False, "flash"), (2, 300, 4, 32, True, "flash"), (1, 1137, 8, 64, False, "tcgen05"),
    (1, 1137, 8, 64, True, "tcgen05"), (1, 4545, 8, 64, False, "tcgen05"), (1, 4545, 8, 64, True, "tcgen05"),
    (1, 4545, 8, 128, False, "tcgen05"), (2, 285, 8, 128, True, "flash"), (2, 256, 2, 64, False, "flash"),
    (1, 257, 2, 64, True, "flash"), (2, 130, 4, 64, True, "flash"), (2, 385, 2, 128, False, "flash"), (1, 700, 2, 16, True, "flash"),
])
def test_attention_kernels_match_torch_fp32(engine, B, N, H, d, causal, impl):
    g = torch.Generator().manual_seed(1000 + N + d)
    qkv = (torch.randn(B, N, 3 * H * d, generator=g) * 1.5).to(torch.bfloat16).to(DEV)
    ref = torch_reference(qkv, H, causal)
    out = engine.attention(qkv, H, causal=causal, impl=impl)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    err = rel_l2(out, ref)
    assert err < 1e-2, (B, N, H, d, causal, impl, err)
    if N <= 1200:  # the fp32-FMA core (strict-mode kernel) on the same inputs
        fma = engine.attention(qkv, H, causal=causal, impl="fma")
        assert rel_l2(fma, ref) < 5e-3


def test_key_tiled_kernel_equals_single_tile_kernel_where_both_apply(engine):
    """<= 256 keys: both tcgen05 kernels see the same bf16 operands; they differ only in the softmax bookkeeping."""
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(2, 200, 3 * 8 * 64, generator=g).to(torch.bfloat16).to(DEV)
    a = engine.attention(qkv, 8, causal=True, impl="tcgen05")
    b = engine.attention(qkv, 8, causal=True, impl="flash")
    assert rel_l2(a, b) < 5e-3


@pytest.mark.parametrize("d,causal", [(64, False), (128, True), (32, False)])
def test_key_tiled_kernel_rescales_when_the_row_maximum_keeps_growing(engine, d, causal):
    """The key-tiled kernel keeps its output accumulator in TMEM and rescales it in place only when a row maximum moves
    by more than 2^8: keys whose magnitude grows along the sequence make the maximum jump in (almost) every key tile, so
    this drives the rescale path (tcgen05.ld / st of the accumulator) on late tiles, not only on the first one."""
    B, N, H = 2, 900, 2
    g = torch.Generator().manual_seed(77 + d)
    qkv = torch.randn(B, N, 3 * H * d, generator=g)
    ramp = torch.linspace(0.5, 6.0, N).view(1, N, 1)
    qkv[:, :, H * d:2 * H * d] *= ramp  # keys
    qkv = qkv.to(torch.bfloat16).to(DEV)
    ref = torch_reference(qkv, H, causal)
    out = engine.attention(qkv, H, causal=causal, impl="flash")
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out, ref) < 1e-2
