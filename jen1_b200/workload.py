"""Algorithmic work of one denoiser evaluation: the roofline numerators of SURVEY.md section 8(d).

`row_work(desc, T)` walks the UNet exactly as reference jen1/model/model.py:225-265 does and counts, for ONE
row (one sample through one UNet pass):

  * `act_elems`  -- every Conv1d / ConvTranspose1d / token-side nn.Linear input read once plus its output
                    written once (elements; x2 bytes in bf16 storage).  GroupNorm / FiLM / SiLU / padding /
                    residual adds / torch.cat are fused away in this engine and therefore add no traffic; the
                    context-side LayerNorm + to_kv of the 128 T5 rows is step-invariant and hoisted.
  * `flops`      -- 2*MAC of every conv / linear / attention contraction on the per-step path.
  * `weight_elems` -- parameters streamed once per step regardless of batch (conv + linear weights, no
                    conditioning networks: those run once per sample() call, not per step).

Used by bench.py (roofline) and DESIGN.md's tables; no torch, no CUDA.
"""
from __future__ import annotations

from dataclasses import dataclass

from .config import UNetDesc


@dataclass
class RowWork:
    act_elems: int = 0
    flops: int = 0
    weight_elems: int = 0
    attn_flops: int = 0

    def conv(self, cin: int, cout: int, k: int, lin: int, lout: int, macs_per_out: int = None):
        self.act_elems += cin * lin + cout * lout
        self.weight_elems += cin * cout * k
        taps = k if macs_per_out is None else macs_per_out
        self.flops += 2 * cin * cout * taps * lout


def row_work(desc: UNetDesc, T: int) -> RowWork:
    w = RowWork()
    H = desc.attention_heads
    M = desc.context_length
    Ls = desc.level_lengths(T)  # Ls[0] = T at to_in's output, Ls[i+1] after down block i

    def resblock(cin, cout, L):
        w.conv(cin, cout, 3, L, L)
        w.conv(cout, cout, 3, L, L)
        if cin != cout:
            w.conv(cin, cout, 1, L, L)

    def transformer(C, N, layers):
        mid = C * desc.attention_multiplier
        w.conv(C, C, 1, N, N)  # conv1x1 in
        for _ in range(layers):
            w.conv(C, 3 * C, 1, N, N)          # self q | k | v
            a = 2 * 2 * N * N * C                # QK^T + PV over all heads
            w.conv(C, C, 1, N, N)              # self to_out
            w.conv(C, C, 1, N, N)              # cross to_q
            a += 2 * 2 * N * M * C
            w.weight_elems += 2 * C * 1          # time-token K/V row comes from the table; cached K/V are read
            w.act_elems += 2 * M * C             # cached cross K/V rows read once per row
            w.conv(C, C, 1, N, N)              # cross to_out
            w.conv(C, mid, 1, N, N)
            w.conv(mid, C, 1, N, N)
            w.flops += a
            w.attn_flops += a
        w.weight_elems -= C * C                  # the second 1x1 conv shares the first one's weights
        w.conv(C, C, 1, N, N)

    c0 = desc.level_channels(0)
    resblock(desc.in_channels + desc.context_channels[0], c0, T)
    n = desc.num_layers
    for i in range(n):
        f = desc.factors[i]
        cin, cout = desc.level_channels(i), desc.level_channels(i + 1)
        w.conv(cin, cout, 2 * f + 1, Ls[i], Ls[i + 1])
        for _ in range(desc.num_blocks[i]):
            resblock(cout, cout, Ls[i + 1])
        if desc.attentions[i] > 0:
            transformer(cout, Ls[i + 1], desc.attentions[i])
    cb = desc.level_channels(n)
    resblock(cb, cb, Ls[n])
    if desc.bottleneck_attention() > 0:
        transformer(cb, Ls[n], desc.bottleneck_attention())
    resblock(cb, cb, Ls[n])
    for i in reversed(range(n)):
        f = desc.factors[i]
        c, cout = desc.level_channels(i + 1), desc.level_channels(i)
        nb = desc.num_blocks[i] + (1 if desc.attentions[i] > 0 else 0)
        for _ in range(nb):
            resblock(2 * c, c, Ls[i + 1])
        if desc.attentions[i] > 0:
            transformer(c, Ls[i + 1], desc.attentions[i])
        if f == 1:
            w.conv(c, cout, 3, Ls[i + 1], Ls[i])
        else:  # ConvTranspose1d k=2f: every output position receives exactly two taps
            w.conv(c, cout, 2 * f, Ls[i + 1], Ls[i], macs_per_out=2)
    resblock(c0, desc.out_channels, T)
    return w


def step_bytes(desc: UNetDesc, B: int, T: int, cfg: bool = True, elem_bytes: int = 2) -> dict:
    """Compulsory HBM bytes of one sampler step on one GPU: weights once + activations per row."""
    rw = row_work(desc, T)
    rows = (2 if cfg else 1) * B
    # sampler state traffic (fp32): x read + noise read + x written + UNet input pack read, per sample
    state = 4 * 4 * desc.in_channels * T * B
    return {
        "weight_bytes": rw.weight_elems * elem_bytes,
        "act_bytes_per_row": rw.act_elems * elem_bytes,
        "rows": rows,
        "state_bytes": state,
        "total_bytes": rw.weight_elems * elem_bytes + rows * rw.act_elems * elem_bytes,
        "flops": rows * rw.flops,
        "attn_flops": rows * rw.attn_flops,
    }
