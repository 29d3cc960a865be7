"""CPU restatement of the Encodec (SEANet) decoder -- TEST INFRASTRUCTURE, never imported by the product package.

The reference decodes latents with `self.audio_encoder.decoder(sample_embs)` (generation.py:130) where
`audio_encoder = EncodecModel.encodec_model_48khz()` (generation.py:34) comes from pip `encodec==0.1.1`, which is NOT
vendored under /root/reference and not installed in this image.  This file restates the published algorithm
(encodec/modules/seanet.py `SEANetDecoder`, encodec/modules/conv.py `SConv1d` / `SConvTranspose1d` / `pad1d` / `unpad1d`,
encodec/modules/lstm.py `SLSTM`, encodec/modules/norm.py) in plain functional PyTorch fp32.

PARITY PIN: the pip package being absent, the restatement is pinned to the Hugging Face port of the same model
(`transformers.models.encodec.modeling_encodec.EncodecDecoder`, present in this image) built with the 48 kHz
configuration and seeded random weights: oracle/make_golden.py `codec` writes tests/golden/codec_decoder.pt and
tests/test_oracle_golden.py checks this file against it.  Against the pip package itself parity is UNPINNED (no
checkpoint and no package offline); tensor names follow the pip package (jen1_b200/codec_config.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _pad_reflect(x, left, right):
    """encodec/modules/conv.py pad1d: reflect padding, with zero-extension first when the signal is too short."""
    length = x.shape[-1]
    max_pad = max(left, right)
    extra = 0
    if length <= max_pad:
        extra = max_pad - length + 1
        x = F.pad(x, (0, extra))
    y = F.pad(x, (left, right), mode="reflect")
    return y[..., : y.shape[-1] - extra] if extra else y


def _gn1(x, w, b, eps):
    return F.group_norm(x, 1, w, b, eps)


def sconv1d(sd, p, x, eps):
    """SConv1d (non-causal, stride 1, dilation 1): reflect pad (k-1) split right = (k-1)//2, left = rest; conv; GroupNorm(1)."""
    w = sd[p + ".conv.conv.weight"]
    k = w.shape[-1]
    total = k - 1
    right = total // 2
    left = total - right
    y = F.conv1d(_pad_reflect(x, left, right), w, sd[p + ".conv.conv.bias"])
    return _gn1(y, sd[p + ".conv.norm.weight"], sd[p + ".conv.norm.bias"], eps)


def sconvtr1d(sd, p, x, stride, eps):
    """SConvTranspose1d (non-causal): transposed conv; GroupNorm(1) over the UNTRIMMED output; trim k - stride samples,
    right = total // 2, left = rest."""
    w = sd[p + ".convtr.convtr.weight"]
    k = w.shape[-1]
    y = F.conv_transpose1d(x, w, sd[p + ".convtr.convtr.bias"], stride=stride)
    y = _gn1(y, sd[p + ".convtr.norm.weight"], sd[p + ".convtr.norm.bias"], eps)
    total = k - stride
    right = total // 2
    left = total - right
    return y[..., left: y.shape[-1] - right]


def slstm(sd, p, x, layers):
    """SLSTM: nn.LSTM over time on [T, B, C] plus the skip connection."""
    seq = x.permute(2, 0, 1)
    inp = seq
    for layer in range(layers):
        w_ih, w_hh = sd["%s.lstm.weight_ih_l%d" % (p, layer)], sd["%s.lstm.weight_hh_l%d" % (p, layer)]
        bias = sd["%s.lstm.bias_ih_l%d" % (p, layer)] + sd["%s.lstm.bias_hh_l%d" % (p, layer)]
        H = w_hh.shape[1]
        h = torch.zeros(inp.shape[1], H, dtype=inp.dtype)
        c = torch.zeros_like(h)
        gx = inp @ w_ih.t() + bias
        outs = []
        for t in range(inp.shape[0]):
            g = gx[t] + h @ w_hh.t()
            i, f, gg, o = g.split(H, dim=1)  # PyTorch gate order: input, forget, cell, output
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        inp = torch.stack(outs, 0)
    return (inp + seq).permute(1, 2, 0)


def resblock(sd, p, x, eps):
    """SEANetResnetBlock (true_skip=False): shortcut conv (k1) of x + [ELU, conv k3, ELU, conv k1](x)."""
    y = sconv1d(sd, p + ".block.1", F.elu(x), eps)
    y = sconv1d(sd, p + ".block.3", F.elu(y), eps)
    return sconv1d(sd, p + ".shortcut", x, eps) + y


def decoder_forward(desc, sd, z, taps=None):
    """z: latent [B, dimension, T] fp32 -> audio [B, channels, T * hop].  `taps` (dict) collects every stage."""
    eps = desc.eps
    x = z
    for idx, kind, cin, cout, k, stride in desc.layers():
        p = "model.%d" % idx
        if kind == "conv":
            x = sconv1d(sd, p, F.elu(x) if idx > 0 else x, eps)
        elif kind == "lstm":
            x = slstm(sd, p, x, desc.lstm_layers)
        elif kind == "convtr":
            x = sconvtr1d(sd, p, F.elu(x), stride, eps)
        else:
            x = resblock(sd, p, x, eps)
        if taps is not None:
            taps[p] = x
    return x


# ------------------------------------------------------------------------------------------------ encoder + RVQ
def sconv1d_strided(sd, p, x, stride, eps):
    """SConv1d with stride (encodec/modules/conv.py): reflect pad left = total - total//2, right = total//2 + the extra
    padding that makes the last frame complete (get_extra_padding_for_conv1d); conv; GroupNorm(1)."""
    import math
    w = sd[p + ".conv.conv.weight"]
    k = w.shape[-1]
    total = k - stride
    length = x.shape[-1]
    n_frames = math.ceil((length - k + total) / stride + 1) - 1
    extra = n_frames * stride + k - total - length
    right = total // 2
    left = total - right
    y = F.conv1d(_pad_reflect(x, left, right + extra), w, sd[p + ".conv.conv.bias"], stride=stride)
    return _gn1(y, sd[p + ".conv.norm.weight"], sd[p + ".conv.norm.bias"], eps)


def encoder_forward(desc, sd, audio, taps=None):
    """audio [N, channels, L] -> latent [N, dimension, ceil(L / hop)] (SEANetEncoder; tensors named encoder.model.N...)."""
    from jen1_b200.codec_config import encoder_layers
    eps = desc.eps
    x = audio
    for idx, kind, cin, cout, k, stride in encoder_layers(desc):
        p = "encoder.model.%d" % idx
        if kind == "conv":
            x = sconv1d(sd, p, F.elu(x) if idx > 0 else x, eps)
        elif kind == "res":
            x = resblock(sd, p, x, eps)
        elif kind == "down":
            x = sconv1d_strided(sd, p, F.elu(x), stride, eps)
        else:
            x = slstm(sd, p, x, desc.lstm_layers)
        if taps is not None:
            taps[p] = x
    return x


def rvq_encode(sd, emb, n_q):
    """ResidualVectorQuantizer.encode (encodec/quantization/core_vq.py): per stage the nearest codebook entry of the
    residual (euclidean; first index on ties), residual -= entry.  emb [N, D, T] -> codes [n_q, N, T]."""
    residual = emb.permute(0, 2, 1)
    codes = []
    for i in range(n_q):
        e = sd["quantizer.vq.layers.%d._codebook.embed" % i]
        flat = residual.reshape(-1, e.shape[1])
        dist = -(flat.pow(2).sum(1, keepdim=True) - 2 * flat @ e.t() + e.pow(2).sum(1)[None])
        ind = dist.max(dim=-1).indices.view(residual.shape[:-1])
        residual = residual - F.embedding(ind, e)
        codes.append(ind)
    return torch.stack(codes)


def rvq_decode(sd, codes):
    """ResidualVectorQuantizer.decode: sum of the selected entries -> [N, D, T] (what reference generation.py:149 returns)."""
    out = 0.0
    for i, ind in enumerate(codes):
        out = out + F.embedding(ind, sd["quantizer.vq.layers.%d._codebook.embed" % i])
    return out.permute(0, 2, 1)
