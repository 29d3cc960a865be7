"""CPU fp32 restatement of the reference denoiser (`UNetCFG1d`) -- TEST INFRASTRUCTURE ONLY.

This module is the parity oracle: a functional, state_dict-driven PyTorch-CPU restatement of the reference's
floating-point algorithm.  It is imported only by tests/, `__graft_entry__.smoke()` and bench.py's
`cpu_baseline` / `--impl reference` legs; the product path (jen1_b200/) never imports it and fails loudly when
its CUDA library is missing.

Pinning: oracle/make_golden.py runs the UNMODIFIED reference (imported from /root/reference in the build
container) on seeded weights/inputs and commits the outputs under tests/golden/; tests/test_oracle_golden.py
checks this restatement against those fixtures (tolerance 2e-5 abs on O(1) outputs: same ATen kernels, only
the op grouping differs).

Every function cites the reference lines it restates (paths relative to the reference root).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------- primitives
def conv1d(x: Tensor, w: Tensor, b: Optional[Tensor], causal: bool, stride: int = 1) -> Tensor:
    """jen1/model/blocks.py:34-53 -- the ctor `padding` is discarded; pad (k-1) on the left when causal,
    else (k-1)//2 on each side, then an unpadded nn.Conv1d."""
    total = w.shape[-1] - 1
    x = F.pad(x, (total, 0)) if causal else F.pad(x, (total // 2, total // 2))
    return F.conv1d(x, w, b, stride=stride)


def conv_block(sd, p: str, x: Tensor, groups: int, scale_shift, causal: bool) -> Tensor:
    """jen1/model/blocks.py:137-145 ConvBlock1d.forward: GN -> FiLM -> SiLU -> conv."""
    x = F.group_norm(x, groups, sd[p + ".groupnorm.weight"], sd[p + ".groupnorm.bias"], eps=1e-5)
    if scale_shift is not None:
        scale, shift = scale_shift
        x = x * (scale + 1) + shift
    x = F.silu(x)
    return conv1d(x, sd[p + ".project.conv.weight"], sd[p + ".project.conv.bias"], causal)


def resnet_block(sd, p: str, x: Tensor, mapping: Tensor, groups: int, causal: bool) -> Tensor:
    """jen1/model/blocks.py:219-231 ResnetBlock1d.forward (+ MappingToScaleShift :161-165)."""
    h = conv_block(sd, p + ".block1", x, groups, None, causal)
    ss = F.linear(F.silu(mapping), sd[p + ".to_scale_shift.to_scale_shift.1.weight"],
                  sd[p + ".to_scale_shift.to_scale_shift.1.bias"]).unsqueeze(-1)
    scale, shift = ss.chunk(2, dim=1)
    h = conv_block(sd, p + ".block2", h, groups, (scale, shift), causal)
    if p + ".to_out.conv.weight" in sd:
        x = conv1d(x, sd[p + ".to_out.conv.weight"], sd[p + ".to_out.conv.bias"], causal)
    return h + x


def attention(sd, p: str, x: Tensor, heads: int, context: Optional[Tensor], context_mask: Optional[Tensor],
              causal: bool) -> Tensor:
    """jen1/model/blocks.py:415-437 Attention.forward + :355-380 AttentionBase.forward (non-flash branch)."""
    ctx = x if context is None else context
    xn = F.layer_norm(x, x.shape[-1:], sd[p + ".norm.weight"], sd[p + ".norm.bias"])
    cn = F.layer_norm(ctx, ctx.shape[-1:], sd[p + ".norm_context.weight"], sd[p + ".norm_context.bias"])
    q = F.linear(xn, sd[p + ".to_q.weight"])
    k, v = F.linear(cn, sd[p + ".to_kv.weight"]).chunk(2, dim=-1)
    if context_mask is not None:  # padded keys: logit 0 and value 0, NOT -inf (blocks.py:431-434)
        m = context_mask.unsqueeze(-1).to(k.dtype)
        k, v = k * m, v * m
    B, N, C = q.shape
    M, d = k.shape[1], C // heads
    qh = q.view(B, N, heads, d).transpose(1, 2)
    kh = k.view(B, M, heads, d).transpose(1, 2)
    vh = v.view(B, M, heads, d).transpose(1, 2)
    sim = torch.matmul(qh, kh.transpose(-1, -2)) * (d ** -0.5)
    if causal:  # blocks.py:315-319 causal_mask: keep j <= i + (M - N)
        keep = ~torch.ones((N, M), dtype=torch.bool, device=sim.device).triu(M - N + 1)
        sim = sim.masked_fill(~keep, -torch.finfo(sim.dtype).max)
    attn = sim.softmax(dim=-1, dtype=torch.float32)
    out = torch.matmul(attn, vh).transpose(1, 2).reshape(B, N, C)
    return F.linear(out, sd[p + ".attention.to_out.weight"], sd[p + ".attention.to_out.bias"])


def transformer1d(sd, p: str, x: Tensor, heads: int, layers: int, context, context_mask, causal: bool) -> Tensor:
    """jen1/model/blocks.py:528-537 Transformer1d.forward, :483-489 TransformerBlock.forward."""
    w, b = sd[p + ".conv1d.conv.weight"], sd[p + ".conv1d.conv.bias"]
    x = F.group_norm(x, 32, sd[p + ".group_norm.weight"], sd[p + ".group_norm.bias"], eps=1e-6)
    x = conv1d(x, w, b, causal).transpose(1, 2)
    for j in range(layers):
        q = "%s.blocks.%d" % (p, j)
        x = attention(sd, q + ".attention", x, heads, None, None, causal) + x
        x = attention(sd, q + ".cross_attention", x, heads, context, context_mask, False) + x
        h = F.gelu(F.linear(x, sd[q + ".feed_forward.0.weight"], sd[q + ".feed_forward.0.bias"]))
        x = F.linear(h, sd[q + ".feed_forward.2.weight"], sd[q + ".feed_forward.2.bias"]) + x
    return conv1d(x.transpose(1, 2), w, b, causal)  # the SAME 1x1 conv a second time


def time_features(t: Tensor, weights: Tensor) -> Tensor:
    """utils/module.py:66-72 LearnedPositionalEmbedding.forward: [t, sin(2 pi t w), cos(2 pi t w)]."""
    x = t.unsqueeze(-1)
    freqs = x * weights.unsqueeze(0) * 2 * math.pi
    return torch.cat((x, torch.cat((freqs.sin(), freqs.cos()), dim=-1)), dim=-1)


def crop_pair(a: Tensor, b: Tensor):
    """utils/module.py:186-204 crop: centre-crop the longer of the two on the last dim."""
    d = a.shape[-1] - b.shape[-1]
    if d == 0:
        return a, b
    # d < 0 (skip longer than x) cannot occur in the UNet: the up-conv yields f*ceil(L/f) >= L frames; the
    # reference's negative-diff slice (module.py:202) is ill-formed, so it is rejected here instead.
    assert d > 0, "skip longer than x"
    start = d // 2
    end = d - start
    return a[:, :, start: a.shape[-1] - end], b


# --------------------------------------------------------------------------------------------- UNet1d
def unet_forward(desc, sd: Dict[str, Tensor], x: Tensor, time: Tensor, *, embedding: Tensor,
                 embedding_mask: Optional[Tensor], channels_list: List[Tensor], causal: bool = False,
                 taps: Optional[dict] = None) -> Tensor:
    """jen1/model/model.py:225-265 UNet1d.forward with :204-223 get_mapping and the block forwards
    blocks.py:617-650 (down), :817-830 (bottleneck), :736-764 (up).  `taps`, when given, receives named
    intermediate activations (used to localise engine bugs in tests)."""
    G, H = desc.resnet_groups, desc.attention_heads
    x = torch.cat([x, channels_list[0]], dim=1)
    tf = time_features(time, sd["to_time.0.0.weights"])
    m = F.gelu(F.linear(tf, sd["to_time.0.1.weight"], sd["to_time.0.1.bias"]))
    m = F.gelu(F.linear(m, sd["to_mapping.0.weight"], sd["to_mapping.0.bias"]))
    mapping = F.gelu(F.linear(m, sd["to_mapping.2.weight"], sd["to_mapping.2.bias"]))
    if taps is not None:
        taps["mapping"] = mapping
    x = resnet_block(sd, "to_in.block", x, mapping, 1, False)  # Patcher: groups=1, never causal (blocks.py:256)
    if taps is not None:
        taps["to_in"] = x
    skips_list = [[x]]
    for i in range(desc.num_layers):
        p = "downsamples.%d" % i
        x = conv1d(x, sd[p + ".downsample.conv.weight"], sd[p + ".downsample.conv.bias"], causal,
                   stride=desc.factors[i])
        skips = []
        for j in range(desc.num_blocks[i]):
            x = resnet_block(sd, "%s.blocks.%d" % (p, j), x, mapping, G, causal)
            skips.append(x)
        if desc.attentions[i] > 0:
            x = transformer1d(sd, p + ".transformer", x, H, desc.attentions[i], embedding, embedding_mask, causal)
            skips.append(x)
        skips_list.append(skips)
        if taps is not None:
            taps["down%d" % i] = x
    x = resnet_block(sd, "bottleneck.pre_block", x, mapping, G, causal)
    if desc.bottleneck_attention() > 0:
        x = transformer1d(sd, "bottleneck.transformer", x, H, desc.bottleneck_attention(), embedding,
                          embedding_mask, causal)
    x = resnet_block(sd, "bottleneck.post_block", x, mapping, G, causal)
    if taps is not None:
        taps["mid"] = x
    skip_scale = 2 ** -0.5 if desc.use_skip_scale else 1.0
    for u, i in enumerate(reversed(range(desc.num_layers))):
        p = "upsamples.%d" % u
        skips = skips_list.pop()
        for j in range(desc.num_blocks[i] + (1 if desc.attentions[i] else 0)):
            a, s = crop_pair(x, skips.pop())
            x = torch.cat([a, s * skip_scale], dim=1)
            x = resnet_block(sd, "%s.blocks.%d" % (p, j), x, mapping, G, causal)
        if desc.attentions[i] > 0:
            x = transformer1d(sd, p + ".transformer", x, H, desc.attentions[i], embedding, embedding_mask, causal)
        f = desc.factors[i]
        if f == 1:  # blocks.py:72-75: plain nn.Conv1d k3 p1, never causal
            x = F.conv1d(x, sd[p + ".upsample.weight"], sd[p + ".upsample.bias"], padding=1)
        else:  # blocks.py:88-95
            x = F.conv_transpose1d(x, sd[p + ".upsample.weight"], sd[p + ".upsample.bias"], stride=f,
                                   padding=f // 2 + f % 2, output_padding=f % 2)
        if taps is not None:
            taps["up%d" % u] = x
    x = x + skips_list.pop()[0]
    if taps is not None:
        taps["pre_out"] = x
    x = resnet_block(sd, "to_out.block", x, mapping, 1, False)  # Unpatcher
    return x


# --------------------------------------------------------------------------------------------- UNetCFG1d
def unet_cfg_forward(desc, sd: Dict[str, Tensor], x: Tensor, time: Tensor, *, embedding: Tensor,
                     embedding_mask: Optional[Tensor] = None, embedding_scale: float = 1.0,
                     embedding_mask_proba: float = 0.0, batch_cfg: bool = False, scale_cfg: bool = False,
                     scale_phi: float = 0.7, channels_list: List[Tensor] = None, causal: bool = False,
                     features=None, drop_mask: Optional[Tensor] = None, taps: Optional[dict] = None) -> Tensor:
    """jen1/model/model.py:299-376 UNetCFG1d.forward.  `drop_mask` (bool [B]) overrides the bernoulli draw of
    utils/module.py:36-42 so tests can fix the cond-dropout pattern; when None the draw is made with the
    global torch RNG exactly where the reference makes it (model.py:325)."""
    b = embedding.shape[0]
    tf = time_features(time, sd["to_time_embedding.0.0.weights"])
    tok = F.gelu(F.linear(tf, sd["to_time_embedding.0.1.weight"], sd["to_time_embedding.0.1.bias"]))
    embedding = torch.cat([embedding, tok.unsqueeze(1)], dim=1)
    if embedding_mask is not None:
        embedding_mask = torch.cat([embedding_mask, torch.ones((b, 1), device=embedding_mask.device)], dim=1)  # bool -> float promotion
    fixed = sd["fixed_embedding.embedding.weight"][: embedding.shape[1]].unsqueeze(0).expand(b, -1, -1)
    if embedding_mask_proba > 0.0:
        if drop_mask is None:
            if embedding_mask_proba == 1:
                drop_mask = torch.ones((b, 1, 1), dtype=torch.bool, device=embedding.device)
            else:
                drop_mask = torch.bernoulli(torch.full((b, 1, 1), embedding_mask_proba, device=embedding.device)).to(torch.bool)
        embedding = torch.where(drop_mask.view(b, 1, 1), fixed, embedding)
    kw = dict(causal=causal, taps=taps)
    if embedding_scale == 1.0:
        return unet_forward(desc, sd, x, time, embedding=embedding, embedding_mask=embedding_mask,
                            channels_list=channels_list, **kw)
    if batch_cfg:
        mask2 = None if embedding_mask is None else torch.cat([embedding_mask, embedding_mask], dim=0)
        both = unet_forward(desc, sd, torch.cat([x, x]), torch.cat([time, time]),
                            embedding=torch.cat([embedding, fixed]), embedding_mask=mask2,
                            channels_list=[torch.cat([c, c]) for c in channels_list], **kw)
        out, out_masked = both.chunk(2, dim=0)
    else:
        out = unet_forward(desc, sd, x, time, embedding=embedding, embedding_mask=embedding_mask,
                           channels_list=channels_list, **kw)
        out_masked = unet_forward(desc, sd, x, time, embedding=fixed, embedding_mask=embedding_mask,
                                  channels_list=channels_list, **kw)
    out_cfg = out_masked + (out - out_masked) * embedding_scale
    if scale_cfg:
        ratio = out.std(dim=1, keepdim=True) / out_cfg.std(dim=1, keepdim=True)
        return scale_phi * (out_cfg * ratio) + (1 - scale_phi) * out_cfg
    return out_cfg


class OracleUNet:
    """Callable with the reference model's call signature (gdm.py:118-125) around the restatement."""

    def __init__(self, desc, sd):
        self.desc, self.sd = desc, sd

    @torch.no_grad()
    def __call__(self, x, time, **kw):
        return unet_cfg_forward(self.desc, self.sd, x, time, **kw)
