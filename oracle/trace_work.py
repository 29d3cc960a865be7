"""Re-derive the roofline numerators of SURVEY.md section 8(d) from the LIVE reference with forward hooks and commit them as
a fixture (tests/golden/work_trace.json) that pins `jen1_b200/workload.py`.

Run in the build container (where /root/reference exists):   python -m oracle.trace_work
TEST INFRASTRUCTURE -- not imported by the product path.

Method: forward hooks on every nn.Conv1d / nn.ConvTranspose1d / nn.Linear of the unmodified `UNetCFG1d` and on
`AttentionBase` (reference jen1/model/blocks.py:322-380), one single-row evaluation (B = 1, embedding_scale = 1 -> one UNet
pass) per latent length.  Per module call: input elements + output elements, 2*MAC, and its parameters (counted once per
module, however often it is called -- Transformer1d calls the same 1x1 conv twice, blocks.py:528-536).  A call is
*step-invariant* when its input does not depend on x: the time / mapping MLPs, the 56 FiLM linears (functions of t only),
and the cross-attention `to_kv` on the context rows; the engine hoists those out of the sampler step
(DESIGN.md section 3), so they are reported separately:

    F_ref  = 2*MAC of everything the reference computes per row           F_alg = F_ref - step-invariant part
    A      = in + out elements of the per-step conv / linear calls         W_step / W_all = parameters streamed per step / all
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import ref_import  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "work_trace.json")
INVARIANT_MARKERS = ("to_time", "to_mapping", "to_scale_shift", "to_features", "cross_attention.to_kv", "cross_attention.norm_context")


def trace(model, T: int, emb_len: int, emb_feat: int, in_ch: int, ctx_ch: int):
    rec = {"F_ref": 0, "F_inv": 0, "A_step": 0, "A_inv": 0, "attn_flops": 0, "calls": 0,
           # bookkeeping differences between the reference's execution and the engine's algorithmic count (workload.py):
           "A_pad": 0,             # zero-padding elements the reference's Conv1d wrapper materialises before nn.Conv1d (blocks.py:44-51)
           "A_selfattn_kv_in": 0,  # self-attention reads its input twice (to_q and to_kv); the engine has one fused q|k|v projection
           "A_cross_kv_out": 0}    # cross-attention K/V rows: computed per step by the reference, READ from the hoisted cache by the engine
    params_all, params_step = {}, {}
    names = {m: n for n, m in model.named_modules()}
    handles = []

    def add(mod, macs, in_e, out_e, pad_e=0):
        name = names[mod]
        inv = any(k in name for k in INVARIANT_MARKERS)
        rec["calls"] += 1
        rec["F_ref"] += 2 * macs
        w = sum(p.numel() for n_, p in mod.named_parameters(recurse=False) if n_ == "weight")
        params_all[name] = w
        if inv:
            rec["F_inv"] += 2 * macs
            rec["A_inv"] += in_e + out_e
        else:
            rec["A_step"] += in_e + out_e
            rec["A_pad"] += pad_e
            params_step[name] = w
            if name.endswith(".attention.to_kv"):
                rec["A_selfattn_kv_in"] += in_e
        if name.endswith("cross_attention.to_kv"):
            rec["A_cross_kv_out"] += out_e

    def conv_hook(mod, inp, out):
        x = inp[0]
        k = mod.kernel_size[0]
        if isinstance(mod, torch.nn.ConvTranspose1d):
            macs = x.shape[0] * mod.in_channels * x.shape[-1] * mod.out_channels * k  # every input feeds k taps
        else:
            macs = out.shape[0] * mod.out_channels * out.shape[-1] * (mod.in_channels // mod.groups) * k
        pad = 0
        if isinstance(mod, torch.nn.Conv1d) and mod.padding[0] == 0:  # the wrapper padded by (k - 1) * dilation before the call
            pad = x.shape[0] * x.shape[1] * (k - 1) * mod.dilation[0]
        add(mod, macs, x.numel(), out.numel(), pad)

    def lin_hook(mod, inp, out):
        x = inp[0]
        add(mod, (x.numel() // mod.in_features) * mod.in_features * mod.out_features, x.numel(), out.numel())

    def attn_hook(mod, inp, out):  # q [B, N, H*d], k / v [B, M, H*d]: QK^T + PV
        q, k = inp[0], inp[1]
        f = 2 * 2 * q.shape[0] * q.shape[1] * k.shape[1] * q.shape[2]
        rec["F_ref"] += f
        rec["attn_flops"] += f

    from jen1.model.blocks import AttentionBase
    for m in model.modules():
        if isinstance(m, (torch.nn.Conv1d, torch.nn.ConvTranspose1d)):
            handles.append(m.register_forward_hook(conv_hook))
        elif isinstance(m, torch.nn.Linear):
            handles.append(m.register_forward_hook(lin_hook))
        elif isinstance(m, AttentionBase):
            handles.append(m.register_forward_hook(attn_hook))
    g = torch.Generator().manual_seed(T)
    x = torch.randn(1, in_ch, T, generator=g)
    t = torch.randint(0, 1000, (1,), generator=g)
    emb = torch.randn(1, emb_len, emb_feat, generator=g)
    mask = torch.ones(1, emb_len, dtype=torch.bool)
    cc = torch.randn(1, ctx_ch, T, generator=g)
    with torch.no_grad():
        model(x, t, embedding=emb, embedding_mask=mask, features=None, channels_list=[cc], embedding_scale=1.0)
    for h in handles:
        h.remove()
    rec["W_all"] = sum(params_all.values())
    rec["W_step"] = sum(params_step.values())
    rec["F_alg"] = rec["F_ref"] - rec["F_inv"]
    return rec


def trace_codec(T: int):
    """The same for the Encodec-48k decoder, on the Hugging Face port of the SEANet decoder (the pip package
    the reference imports is absent offline): 2*MAC of every Conv1d / ConvTranspose1d / LSTM call for one sample of T frames."""
    from transformers import EncodecConfig
    from transformers.models.encodec.modeling_encodec import EncodecDecoder
    from jen1_b200.codec_config import CodecDesc
    desc = CodecDesc()
    cfg = EncodecConfig(sampling_rate=48000, audio_channels=desc.channels, normalize=True, chunk_length_s=1.0, overlap=0.01,
                        hidden_size=desc.dimension, num_filters=desc.n_filters, num_residual_layers=1,
                        upsampling_ratios=list(desc.ratios), norm_type="time_group_norm", kernel_size=desc.kernel_size,
                        last_kernel_size=desc.last_kernel_size, residual_kernel_size=desc.residual_kernel_size,
                        dilation_growth_rate=2, use_causal_conv=False, pad_mode="reflect", compress=desc.compress,
                        num_lstm_layers=desc.lstm_layers, trim_right_ratio=1.0, use_conv_shortcut=True)
    dec = EncodecDecoder(cfg).eval()
    rec = {"flops": 0, "conv_out_elems": 0, "samples": 0}

    def conv_hook(mod, inp, out):
        x = inp[0]
        k = mod.kernel_size[0]
        if isinstance(mod, torch.nn.ConvTranspose1d):
            rec["flops"] += 2 * x.shape[1] * x.shape[-1] * mod.out_channels * k
        else:
            rec["flops"] += 2 * mod.out_channels * out.shape[-1] * mod.in_channels * k
        rec["conv_out_elems"] += out.numel()

    def lstm_hook(mod, inp, out):
        L = inp[0].shape[0] if not mod.batch_first else inp[0].shape[1]
        for layer in range(mod.num_layers):
            cin = mod.input_size if layer == 0 else mod.hidden_size
            rec["flops"] += 2 * L * 4 * mod.hidden_size * (cin + mod.hidden_size)

    for m in dec.modules():
        if isinstance(m, (torch.nn.Conv1d, torch.nn.ConvTranspose1d)):
            m.register_forward_hook(conv_hook)
        elif isinstance(m, torch.nn.LSTM):
            m.register_forward_hook(lstm_hook)
    with torch.no_grad():
        y = dec(torch.randn(1, desc.dimension, T, generator=torch.Generator().manual_seed(T)))
    rec["samples"] = y.shape[-1]
    return rec


def main():
    torch.manual_seed(0)
    model = ref_import.build_reference_unet().eval()
    kw = ref_import.reference_model_kwargs()
    out = {"what": "forward-hook trace of the unmodified reference UNetCFG1d, one row (B = 1, one UNet pass), per latent length T; "
                   "generated by oracle/trace_work.py", "per_T": {}}
    for T in (150, 1515, 4545):
        out["per_T"][str(T)] = trace(model, T, kw["context_embedding_max_length"], kw["context_embedding_features"],
                                     kw["in_channels"], kw["context_channels"][0])
        print(T, out["per_T"][str(T)])
    out["codec_decoder_per_T"] = {}
    for T in (150, 600):
        out["codec_decoder_per_T"][str(T)] = trace_codec(T)
        print("codec", T, out["codec_decoder_per_T"][str(T)])
    with open(GOLD, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", GOLD)


if __name__ == "__main__":
    main()
