"""Model / diffusion hyper-parameters and the state_dict tensor inventory of the JEN-1 denoiser.

Values mirror the reference's class-attribute "dataclasses" (reference utils/config.py:49-74 `ModelConfig`,
:23-33 `GDM_Config`); the mechanism (introspecting `__dict__` of a class used as a namespace,
utils/script_util.py:275) is replaced by one plain dataclass.  `UNetDesc.tensor_spec()` enumerates every
parameter of `UNetCFG1d` with the exact key names / shapes / order the reference's `state_dict()` produces
(reference jen1/model/model.py:14-181, 271-297; jen1/model/blocks.py) -- that list is part of the drop-in
boundary (SURVEY.md section 8b "Weights") and is checked against the live reference by
tests/test_oracle_golden.py through a committed fixture.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple


@dataclass
class UNetDesc:
    # reference utils/config.py:49-74
    in_channels: int = 128
    channels: int = 128
    multipliers: Sequence[int] = (1, 1, 1, 2, 2, 4, 4, 4, 8, 8)
    factors: Sequence[int] = (1, 4, 4, 4, 2, 2, 2, 2, 2)
    num_blocks: Sequence[int] = (1, 3, 3, 3, 3, 3, 3, 3, 1)
    attentions: Sequence[int] = (0, 0, 0, 1, 1, 1, 1, 1, 1)
    patch_size: int = 1
    resnet_groups: int = 8
    use_context_time: bool = True
    kernel_multiplier_downsample: int = 2
    use_nearest_upsample: bool = False
    use_skip_scale: bool = True
    use_xattn_time: bool = True
    out_channels: int = 128
    context_features_multiplier: int = 4
    context_channels: Sequence[int] = (129,)
    context_embedding_features: int = 1024
    context_embedding_max_length: int = 128
    attention_heads: int = 8
    attention_multiplier: int = 1

    def __post_init__(self):
        self.multipliers = tuple(self.multipliers)
        self.factors = tuple(self.factors)
        self.num_blocks = tuple(self.num_blocks)
        self.attentions = tuple(self.attentions)
        self.context_channels = tuple(self.context_channels)
        n = self.num_layers
        assert len(self.factors) == n and len(self.num_blocks) == n and len(self.attentions) >= n
        # Features this build does not re-host (dead or broken in the reference).
        assert self.patch_size == 1, "patch_size != 1 is not supported (reference default is 1)"
        assert self.kernel_multiplier_downsample == 2
        assert not self.use_nearest_upsample
        assert self.use_context_time and self.use_xattn_time
        assert len(self.context_channels) == 1, "only level-0 input-concat conditioning is supported"

    # ---- derived -----------------------------------------------------------------------------------
    @property
    def num_layers(self) -> int:
        return len(self.multipliers) - 1

    @property
    def mapping_features(self) -> int:  # reference model.py:73
        return self.channels * self.context_features_multiplier

    @property
    def time_dim(self) -> int:  # LearnedPositionalEmbedding(dim=channels): dim+1 features
        return self.channels + 1

    @property
    def context_length(self) -> int:  # +1 time token, reference model.py:293
        return self.context_embedding_max_length + 1

    def level_channels(self, i: int) -> int:
        return self.channels * self.multipliers[i]

    def level_lengths(self, T: int) -> List[int]:
        """Sequence length at the output of to_in and of each down block."""
        out = [T]
        for f in self.factors:
            out.append(-(-out[-1] // f))
        return out

    def bottleneck_attention(self) -> int:
        return self.attentions[-1]

    # ---- state_dict inventory ----------------------------------------------------------------------
    def tensor_spec(self) -> List[Tuple[str, Tuple[int, ...], str]]:
        """(name, shape, kind) for all parameters, in the reference's registration order."""
        S: List[Tuple[str, Tuple[int, ...], str]] = []
        Fm = self.mapping_features
        E = self.context_embedding_features

        def lin(p, o, i, bias=True):
            S.append((p + ".weight", (o, i), "linear_w:%d" % i))
            if bias:
                S.append((p + ".bias", (o,), "bias:%d" % i))

        def norm(p, c):
            S.append((p + ".weight", (c,), "norm_w"))
            S.append((p + ".bias", (c,), "norm_b"))

        def conv(p, o, i, k, wrapped=True):
            q = p + (".conv" if wrapped else "")
            S.append((q + ".weight", (o, i, k), "conv_w:%d" % (i * k)))
            S.append((q + ".bias", (o,), "bias:%d" % (i * k)))

        def resblock(p, cin, cout):
            norm(p + ".block1.groupnorm", cin)
            conv(p + ".block1.project", cout, cin, 3)
            lin(p + ".to_scale_shift.to_scale_shift.1", 2 * cout, Fm)
            norm(p + ".block2.groupnorm", cout)
            conv(p + ".block2.project", cout, cout, 3)
            if cin != cout:
                conv(p + ".to_out", cout, cin, 1)

        def attention(p, c, ctx):
            norm(p + ".norm", c)
            norm(p + ".norm_context", ctx)
            lin(p + ".to_q", c, c, bias=False)
            lin(p + ".to_kv", 2 * c, ctx, bias=False)
            lin(p + ".attention.to_out", c, c)

        def transformer(p, c, layers):
            norm(p + ".group_norm", c)
            conv(p + ".conv1d", c, c, 1)
            for j in range(layers):
                q = "%s.blocks.%d" % (p, j)
                attention(q + ".attention", c, c)
                attention(q + ".cross_attention", c, E)
                lin(q + ".feed_forward.0", c * self.attention_multiplier, c)
                lin(q + ".feed_forward.2", c, c * self.attention_multiplier)

        # UNet1d.__init__ registration order: to_mapping, to_time, to_in, downsamples, bottleneck, upsamples, to_out
        lin("to_mapping.0", Fm, Fm)
        lin("to_mapping.2", Fm, Fm)
        S.append(("to_time.0.0.weights", (self.channels // 2,), "posemb"))
        lin("to_time.0.1", Fm, self.time_dim)
        resblock("to_in.block", self.in_channels + self.context_channels[0], self.level_channels(0))
        for i in range(self.num_layers):
            p = "downsamples.%d" % i
            cin, cout, f = self.level_channels(i), self.level_channels(i + 1), self.factors[i]
            conv(p + ".downsample", cout, cin, 2 * f + 1)
            for j in range(self.num_blocks[i]):
                resblock("%s.blocks.%d" % (p, j), cout, cout)
            if self.attentions[i] > 0:
                transformer(p + ".transformer", cout, self.attentions[i])
        cb = self.level_channels(self.num_layers)
        resblock("bottleneck.pre_block", cb, cb)
        if self.bottleneck_attention() > 0:
            transformer("bottleneck.transformer", cb, self.bottleneck_attention())
        resblock("bottleneck.post_block", cb, cb)
        for u, i in enumerate(reversed(range(self.num_layers))):
            p = "upsamples.%d" % u
            cin, cout, f = self.level_channels(i + 1), self.level_channels(i), self.factors[i]
            for j in range(self.num_blocks[i] + (1 if self.attentions[i] else 0)):
                resblock("%s.blocks.%d" % (p, j), 2 * cin, cin)
            if self.attentions[i] > 0:
                transformer(p + ".transformer", cin, self.attentions[i])
            if f == 1:
                conv(p + ".upsample", cout, cin, 3, wrapped=False)
            else:
                S.append((p + ".upsample.weight", (cin, cout, 2 * f), "conv_w:%d" % (cout * 2 * f)))
                S.append((p + ".upsample.bias", (cout,), "bias:%d" % (cout * 2 * f)))
        resblock("to_out.block", self.level_channels(0), self.out_channels)
        # UNetCFG1d additions
        S.append(("to_time_embedding.0.0.weights", (self.channels // 2,), "posemb"))
        lin("to_time_embedding.0.1", E, self.time_dim)
        S.append(("fixed_embedding.embedding.weight", (self.context_length, E), "embedding"))
        return S

    def num_params(self) -> int:
        n = 0
        for _, shape, _ in self.tensor_spec():
            m = 1
            for s in shape:
                m *= s
            n += m
        return n


def tiny_desc() -> UNetDesc:
    """A reduced architecture with every structural feature of the full model (strided / transposed convs
    with factors 1, 2 and 4, skip crops, self/cross attention, bottleneck transformer) used for fast tests."""
    return UNetDesc(in_channels=8, channels=32, multipliers=(1, 1, 2, 2, 4), factors=(1, 4, 2, 2),
                    num_blocks=(1, 2, 2, 1), attentions=(0, 0, 1, 1), out_channels=8, context_channels=(9,),
                    context_embedding_features=64, context_embedding_max_length=12, attention_heads=4)


@dataclass
class DiffusionDesc:
    # reference utils/config.py:23-33
    steps: int = 1000
    noise_schedule: str = "linear"
    objective: str = "noise"
    loss_type: str = "l2"
    cfg_dropout_proba: float = 0.2
    embedding_scale: float = 0.8
    batch_cfg: bool = True
    scale_cfg: bool = True
    ddim_sampling_eta: float = 1.0
    scale_phi: float = 0.7  # reference model.py:310 default


def latent_frames(seconds: float, sample_rate: int = 48000, hop: int = 320, segment: int = 48000,
                  stride: int = 47520) -> int:
    """Latent length the Encodec-48k front end produces (reference generation.py:145-150):
    1 s segments at stride 47 520 samples, ceil(len/320) frames each, concatenated."""
    total = int(round(seconds * sample_rate))
    frames, off = 0, 0
    while off < total:
        seg = min(segment, total - off)
        frames += -(-seg // hop)
        off += stride
    return frames
