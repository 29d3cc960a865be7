"""Architecture description of the Encodec (SEANet) decoder JEN-1 decodes its latents with.

Reference: generation.py:34 builds `EncodecModel.encodec_model_48khz()` (pip `encodec==0.1.1`, not vendored in the
reference tree) and generation.py:130 calls its `decoder` module directly on the sampled latent `[B, 128, T]`.
The 48 kHz model's decoder is `SEANetDecoder(channels=2, dimension=128, n_filters=32, n_residual_layers=1,
ratios=[8, 5, 4, 2], activation='ELU', norm='time_group_norm', kernel_size=7, last_kernel_size=7,
residual_kernel_size=3, dilation_base=2, causal=False, pad_mode='reflect', true_skip=False, compress=2, lstm=2)`
(encodec/model.py `encodec_model_48khz`, encodec/modules/seanet.py `SEANetDecoder`):

    model.0   SConv1d(dimension -> 16*n_filters, k7)             every conv: reflect pad, then GroupNorm(1, Cout)
    model.1   SLSTM(16*n_filters, 2 layers) + skip
    per ratio r (channels C -> C/2):
      ELU, SConvTranspose1d(C -> C/2, k=2r, stride r)           GroupNorm over the untrimmed output, then trim r/2 | r - r/2
      SEANetResnetBlock(C/2): [ELU, SConv1d(C/2 -> C/4, k3), ELU, SConv1d(C/4 -> C/2, k1)] + SConv1d shortcut (k1)
    ELU, SConv1d(n_filters -> channels, k7)

Tensor names: the canonical names used by this package are the pip package's (`model.N.conv.conv.weight`,
`model.N.conv.norm.weight`, `model.N.convtr.convtr.weight`, `model.N.lstm.weight_ih_l0`, `model.N.block.K...`,
`model.N.shortcut...`); `canonical_state_dict` also accepts the Hugging Face port's layout (`layers.N.conv.weight`,
`layers.N.norm.weight`, ...) and an optional `decoder.` prefix.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch


@dataclass(frozen=True)
class CodecDesc:
    channels: int = 2
    dimension: int = 128
    n_filters: int = 32
    ratios: Tuple[int, ...] = (8, 5, 4, 2)
    kernel_size: int = 7
    last_kernel_size: int = 7
    residual_kernel_size: int = 3
    compress: int = 2
    lstm_layers: int = 2
    eps: float = 1e-5

    @property
    def hidden(self) -> int:
        return self.n_filters * (2 ** len(self.ratios))

    @property
    def hop(self) -> int:
        return int(math.prod(self.ratios))

    def layers(self) -> List[tuple]:
        """[(index, kind, cin, cout, k, stride)] in `model.N` order (ELU entries are skipped, they hold no tensors)."""
        out = [(0, "conv", self.dimension, self.hidden, self.kernel_size, 1), (1, "lstm", self.hidden, self.hidden, 0, 0)]
        idx, c = 2, self.hidden
        for r in self.ratios:
            out.append((idx + 1, "convtr", c, c // 2, 2 * r, r))
            out.append((idx + 2, "res", c // 2, c // 2, self.residual_kernel_size, 1))
            idx += 3
            c //= 2
        out.append((idx + 1, "conv", c, self.channels, self.last_kernel_size, 1))
        return out

    def tensor_spec(self) -> List[tuple]:
        """[(name, shape, kind)] -- every tensor of the decoder's state_dict, pip-package naming."""
        spec = []

        def conv(prefix, cin, cout, k):
            spec.append((prefix + ".conv.conv.weight", (cout, cin, k), "w:%d" % (cin * k)))
            spec.append((prefix + ".conv.conv.bias", (cout,), "w:%d" % (cin * k)))
            spec.append((prefix + ".conv.norm.weight", (cout,), "norm_w"))
            spec.append((prefix + ".conv.norm.bias", (cout,), "norm_b"))

        for idx, kind, cin, cout, k, stride in self.layers():
            p = "model.%d" % idx
            if kind == "conv":
                conv(p, cin, cout, k)
            elif kind == "lstm":
                for layer in range(self.lstm_layers):
                    spec.append(("%s.lstm.weight_ih_l%d" % (p, layer), (4 * cout, cin), "w:%d" % cout))
                    spec.append(("%s.lstm.weight_hh_l%d" % (p, layer), (4 * cout, cout), "w:%d" % cout))
                    spec.append(("%s.lstm.bias_ih_l%d" % (p, layer), (4 * cout,), "w:%d" % cout))
                    spec.append(("%s.lstm.bias_hh_l%d" % (p, layer), (4 * cout,), "w:%d" % cout))
            elif kind == "convtr":
                spec.append((p + ".convtr.convtr.weight", (cin, cout, k), "w:%d" % (cin * 2)))
                spec.append((p + ".convtr.convtr.bias", (cout,), "w:%d" % (cin * 2)))
                spec.append((p + ".convtr.norm.weight", (cout,), "norm_w"))
                spec.append((p + ".convtr.norm.bias", (cout,), "norm_b"))
            else:
                hid = cin // self.compress
                conv(p + ".block.1", cin, hid, k)
                conv(p + ".block.3", hid, cout, 1)
                conv(p + ".shortcut", cin, cout, 1)
        return spec


def encoder_layers(desc: CodecDesc) -> List[tuple]:
    """SEANetEncoder (encodec/modules/seanet.py), the decoder's mirror: [(index, kind, cin, cout, k, stride)] in `model.N`
    order: conv k7 channels -> n_filters; per ratio r (reversed decoder ratios, channels C -> 2C): ResnetBlock(C), ELU,
    SConv1d(C -> 2C, k = 2r, stride r); SLSTM; ELU; SConv1d(16*n_filters -> dimension, k7)."""
    out = [(0, "conv", desc.channels, desc.n_filters, desc.kernel_size, 1)]
    idx, c = 0, desc.n_filters
    for r in reversed(desc.ratios):
        out.append((idx + 1, "res", c, c, desc.residual_kernel_size, 1))
        out.append((idx + 3, "down", c, 2 * c, 2 * r, r))
        idx += 3
        c *= 2
    out.append((idx + 1, "lstm", c, c, 0, 0))
    out.append((idx + 3, "conv", c, desc.dimension, desc.last_kernel_size, 1))
    return out


def encoder_tensor_spec(desc: CodecDesc, n_q: int = 16, codebook_size: int = 1024) -> List[tuple]:
    """[(name, shape, kind)]: the encoder's tensors (`encoder.model.N...`, pip naming) and the residual vector quantizer's
    codebooks (`quantizer.vq.layers.i._codebook.embed`)."""
    spec = []

    def conv(prefix, cin, cout, k):
        spec.append((prefix + ".conv.conv.weight", (cout, cin, k), "w:%d" % (cin * k)))
        spec.append((prefix + ".conv.conv.bias", (cout,), "w:%d" % (cin * k)))
        spec.append((prefix + ".conv.norm.weight", (cout,), "norm_w"))
        spec.append((prefix + ".conv.norm.bias", (cout,), "norm_b"))

    for idx, kind, cin, cout, k, stride in encoder_layers(desc):
        p = "encoder.model.%d" % idx
        if kind in ("conv", "down"):
            conv(p, cin, cout, k)
        elif kind == "lstm":
            for layer in range(desc.lstm_layers):
                spec.append(("%s.lstm.weight_ih_l%d" % (p, layer), (4 * cout, cin), "w:%d" % cout))
                spec.append(("%s.lstm.weight_hh_l%d" % (p, layer), (4 * cout, cout), "w:%d" % cout))
                spec.append(("%s.lstm.bias_ih_l%d" % (p, layer), (4 * cout,), "w:%d" % cout))
                spec.append(("%s.lstm.bias_hh_l%d" % (p, layer), (4 * cout,), "w:%d" % cout))
        else:
            hid = cin // desc.compress
            conv(p + ".block.1", cin, hid, k)
            conv(p + ".block.3", hid, cout, 1)
            conv(p + ".shortcut", cin, cout, 1)
    for i in range(n_q):
        spec.append(("quantizer.vq.layers.%d._codebook.embed" % i, (codebook_size, desc.dimension), "codebook"))
    return spec


def random_encoder_state_dict(desc: CodecDesc, seed: int = 0, n_q: int = 16, codebook_size: int = 1024) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    sd = {}
    for name, shape, kind in encoder_tensor_spec(desc, n_q, codebook_size):
        if kind.startswith("w:"):
            bound = 1.0 / math.sqrt(int(kind[2:]))
            t = (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound
        elif kind == "norm_w":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
        elif kind == "norm_b":
            t = 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
        else:  # codebooks: stage i has residual-sized entries
            i = int(name.split(".")[3])
            t = torch.randn(shape, generator=g, dtype=torch.float32) * (0.7 ** i)
        sd[name] = t
    return sd


def canonical_encoder_state_dict(desc: CodecDesc, sd: Dict[str, torch.Tensor], n_q: int = 16, codebook_size: int = 1024):
    """pip layout (`encoder.model.N.conv.conv.weight`, `quantizer.vq.layers.i._codebook.embed`) or the Hugging Face
    port's (`encoder.layers.N.conv.weight`, `quantizer.layers.i.codebook.embed`) -> canonical (pip) names."""
    out = {}
    for name, shape, _ in encoder_tensor_spec(desc, n_q, codebook_size):
        hf = name.replace("encoder.model.", "encoder.layers.", 1)
        hf = hf.replace(".conv.conv.", ".conv.").replace(".conv.norm.", ".norm.")
        hf = hf.replace("quantizer.vq.layers.", "quantizer.layers.").replace("._codebook.embed", ".codebook.embed")
        for c in (name, hf):
            if c in sd:
                t = sd[c].detach().to(torch.float32).cpu().contiguous()
                if tuple(t.shape) != tuple(shape):
                    raise ValueError("tensor %s has shape %s, expected %s" % (c, tuple(t.shape), tuple(shape)))
                out[name] = t
                break
        else:
            raise KeyError("encoder state_dict is missing %s (also tried %s)" % (name, hf))
    return out


def decode_work(desc: CodecDesc, T: int) -> Dict[str, float]:
    """Algorithmic work of ONE sample's decode at T latent frames: fp32 bytes every layer must read (each operand once;
    a "sum of two normalised tensors" input counts both) and write, and 2*MAC flops.  Used by bench.py's codec leg."""
    by = fl = 0.0
    L, two = T, False  # `two`: the running value is a sum of two stored tensors
    for idx, kind, cin, cout, k, stride in desc.layers():
        nin = 2 if two else 1
        if kind == "conv":
            by += 4.0 * L * (nin * cin + cout)
            fl += 2.0 * L * cin * cout * k
            two = False
        elif kind == "lstm":
            for _ in range(desc.lstm_layers):
                by += 4.0 * L * (cin + 4 * cin) + 4.0 * L * (4 * cin + cin)  # projection in/out, recurrence in/out
                fl += 2.0 * L * cin * 4 * cin * 2
            two = True  # lstm output + skip
        elif kind == "convtr":
            by += 4.0 * (L * nin * cin + (L + 1) * stride * cout)
            fl += 2.0 * L * cin * cout * k
            L *= stride
            two = False
        else:
            hid = cin // desc.compress
            by += 4.0 * L * (cin + hid) + 4.0 * L * (hid + cout) + 4.0 * L * (cin + cout)
            fl += 2.0 * L * (cin * hid * k + hid * cout + cin * cout)
            two = True
    by += 4.0 * L * desc.channels * 2  # final GroupNorm pass: read raw, write audio
    return {"bytes": by, "flops": fl, "samples": float(L)}


def tiny_codec_desc() -> CodecDesc:
    return CodecDesc(channels=2, dimension=16, n_filters=4, ratios=(4, 2), kernel_size=7, last_kernel_size=7)


def random_state_dict(desc: CodecDesc, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded random-init weights (PyTorch-default-like scales, non-trivial norm affines); the pip package's
    checkpoint cannot be fetched offline."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    sd = {}
    for name, shape, kind in desc.tensor_spec():
        if kind.startswith("w:"):
            bound = 1.0 / math.sqrt(int(kind[2:]))
            t = (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound
        elif kind == "norm_w":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
        else:
            t = 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
        sd[name] = t
    return sd


def canonical_state_dict(desc: CodecDesc, sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Normalise a decoder state_dict (pip `encodec` layout, or the Hugging Face port's, with or without a `decoder.`
    prefix; extra keys such as the encoder's / quantizer's are ignored) to the canonical names; checks completeness."""
    src = {}
    for k, v in sd.items():
        if k.startswith("decoder."):
            k = k[len("decoder."):]
        src[k] = v
    out = {}
    for name, shape, _ in desc.tensor_spec():
        cands = [name]
        hf = name.replace("model.", "layers.", 1)
        hf = hf.replace(".conv.conv.", ".conv.").replace(".conv.norm.", ".norm.")
        hf = hf.replace(".convtr.convtr.", ".conv.").replace(".convtr.norm.", ".norm.")
        cands.append(hf)
        for c in cands:
            if c in src:
                t = src[c].detach().to(torch.float32).cpu().contiguous()
                if tuple(t.shape) != tuple(shape):
                    raise ValueError("tensor %s has shape %s, expected %s" % (c, tuple(t.shape), tuple(shape)))
                out[name] = t
                break
        else:
            raise KeyError("decoder state_dict is missing %s (also tried %s)" % (name, hf))
    return out


def to_hf_names(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Canonical names -> the Hugging Face `EncodecDecoder` layout (used by oracle/make_golden.py)."""
    out = {}
    for name, t in sd.items():
        hf = name.replace("model.", "layers.", 1)
        hf = hf.replace(".conv.conv.", ".conv.").replace(".conv.norm.", ".norm.")
        hf = hf.replace(".convtr.convtr.", ".conv.").replace(".convtr.norm.", ".norm.")
        out[hf] = t
    return out
