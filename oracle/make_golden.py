"""Generate the committed golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):   python -m oracle.make_golden
TEST INFRASTRUCTURE -- not imported by the product path.  The reference ships no tests, checkpoints or golden
vectors, so the pins are outputs of the live reference on seeded weights
(`jen1_b200.weights.random_state_dict`, loaded with the reference's own `load_state_dict`) and seeded inputs.
Weights are NOT stored (1.2 GB): fixtures hold seeds, inputs and reference outputs only.
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

from jen1_b200.config import UNetDesc, tiny_desc  # noqa: E402
from jen1_b200.weights import random_state_dict  # noqa: E402
from oracle import ref_import  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def desc_overrides(desc: UNetDesc) -> dict:
    return dict(in_channels=desc.in_channels, channels=desc.channels, multipliers=list(desc.multipliers),
                factors=list(desc.factors), num_blocks=list(desc.num_blocks), attentions=list(desc.attentions),
                out_channels=desc.out_channels, context_channels=list(desc.context_channels),
                context_embedding_features=desc.context_embedding_features,
                context_embedding_max_length=desc.context_embedding_max_length,
                attention_heads=desc.attention_heads, attention_multiplier=desc.attention_multiplier,
                resnet_groups=desc.resnet_groups)


def make_inputs(desc: UNetDesc, B: int, T: int, seed: int, masked_tail: int = 0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, desc.in_channels, T, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    emb = torch.randn(B, desc.context_embedding_max_length, desc.context_embedding_features, generator=g)
    mask = torch.ones(B, desc.context_embedding_max_length, dtype=torch.bool)
    if masked_tail:
        mask[:, -masked_tail:] = False
        emb = emb * mask.unsqueeze(-1)  # T5Conditioner zeroes padded rows (conditioners.py:109)
    cc = torch.randn(B, desc.context_channels[0], T, generator=g)
    return x, t, emb, mask, cc


VARIANTS = {
    "plain": dict(embedding_scale=1.0),
    "cfg": dict(embedding_scale=0.8, batch_cfg=True, scale_cfg=True, embedding_mask_proba=0.0),
    "cfg_causal": dict(embedding_scale=0.8, batch_cfg=True, scale_cfg=True, embedding_mask_proba=0.0, causal=True),
    "cfg_noscale": dict(embedding_scale=0.8, batch_cfg=True, scale_cfg=False, embedding_mask_proba=0.0),
}


def run_unet_cases(desc, model, cases):
    out = {}
    with torch.no_grad():
        for name, (B, T, seed, masked_tail, variants) in cases.items():
            x, t, emb, mask, cc = make_inputs(desc, B, T, seed, masked_tail)
            rec = dict(B=B, T=T, seed=seed, masked_tail=masked_tail, outputs={})
            for v in variants:
                kw = dict(VARIANTS[v])
                y = model(x, t, embedding=emb, embedding_mask=mask, features=None, channels_list=[cc], **kw)
                rec["outputs"][v] = y.clone()
            # one stochastic cond-dropout case: the reference draws bernoulli from the global RNG (model.py:325)
            torch.manual_seed(1234 + seed)
            y = model(x, t, embedding=emb, embedding_mask=mask, features=None, channels_list=[cc],
                      embedding_scale=0.8, batch_cfg=True, scale_cfg=True, embedding_mask_proba=0.5)
            rec["outputs"]["cfg_dropout_p0.5_seed%d" % (1234 + seed)] = y.clone()
            out[name] = rec
    return out


def make_config2_golden():
    """Full-size UNet at BASELINE config 2's shape (B=1, T=1515, CFG): pins the benchmarked configuration to the
    live reference, not only to the port.  Kept in its own file so the round-1 fixtures stay byte-identical."""
    ref_import.install_shims()
    desc = UNetDesc()
    sd = random_state_dict(desc, seed=0)
    model = ref_import.build_reference_unet()
    model.load_state_dict(sd, strict=True)
    cases = {"T1515_B1": (1, 1515, 201, 0, ["cfg"])}
    full = run_unet_cases(desc, model, cases)
    torch.save(dict(weights_seed=0, cases=full), os.path.join(GOLD, "unet_full_c2.pt"))
    print("  %-28s %8.1f KB" % ("unet_full_c2.pt", os.path.getsize(os.path.join(GOLD, "unet_full_c2.pt")) / 1024))


def make_conditioner_golden():
    """Int / Number conditioners of the reference (jen1/conditioners.py:114-164) on seeded weights: state dicts, inputs
    and outputs (the T5 conditioner's pretrained weights are unreachable offline: parity unpinned for it)."""
    ref_import.install_shims()
    from jen1.conditioners import IntConditioner, NumberConditioner
    torch.manual_seed(3)
    ic, nc = IntConditioner(64, 0, 512), NumberConditioner(64, 0, 512)
    ints, floats = [3, 600, 0, 77], [3.0, 700.0, 100.5, 0.25]
    with torch.no_grad():
        io, no = ic(ints, "cpu"), nc(floats, "cpu")
    torch.save(dict(int_sd=ic.state_dict(), num_sd=nc.state_dict(), ints=ints, floats=floats,
                    int_out=[t.clone() for t in io], num_out=[t.clone() for t in no]),
               os.path.join(GOLD, "conditioners.pt"))
    print("  %-28s %8.1f KB" % ("conditioners.pt", os.path.getsize(os.path.join(GOLD, "conditioners.pt")) / 1024))


def make_codec_golden():
    """Encodec-48k decoder: the pip package the reference imports (generation.py:9,34) is absent offline; the pin is the
    Hugging Face port of the same SEANet decoder, 48 kHz configuration, seeded random weights (never stored: they are
    regenerated from the seed by jen1_b200.codec_config.random_state_dict)."""
    from transformers import EncodecConfig
    from transformers.models.encodec.modeling_encodec import EncodecDecoder
    from jen1_b200.codec_config import CodecDesc, random_state_dict as codec_sd, to_hf_names
    desc = CodecDesc()
    cfg = EncodecConfig(sampling_rate=48000, audio_channels=desc.channels, normalize=True, chunk_length_s=1.0, overlap=0.01,
                        hidden_size=desc.dimension, num_filters=desc.n_filters, num_residual_layers=1,
                        upsampling_ratios=list(desc.ratios), norm_type="time_group_norm", kernel_size=desc.kernel_size,
                        last_kernel_size=desc.last_kernel_size, residual_kernel_size=desc.residual_kernel_size,
                        dilation_growth_rate=2, use_causal_conv=False, pad_mode="reflect", compress=desc.compress,
                        num_lstm_layers=desc.lstm_layers, trim_right_ratio=1.0, use_conv_shortcut=True)
    dec = EncodecDecoder(cfg).eval()
    sd = codec_sd(desc, 11)
    missing, unexpected = dec.load_state_dict(to_hf_names(sd), strict=True)
    cases = {}
    for name, B, T, seed in (("b2_t20", 2, 20, 5), ("b1_t3", 1, 3, 6), ("b3_t33", 3, 33, 7)):
        g = torch.Generator().manual_seed(seed)
        z = torch.randn(B, desc.dimension, T, generator=g)
        with torch.no_grad():
            taps, x = {}, z
            for i, layer in enumerate(dec.layers):
                x = layer(x)
                if name == "b2_t20" and i in (0, 1, 3, 4, 15):
                    taps["model.%d" % i] = x[:, :, :32].clone()  # head of every stage kind (conv, lstm, convtr, res, last)
            cases[name] = dict(z=z, out=x.clone(), taps=taps)
        print("  codec %-8s out %s  |out| max %.3f" % (name, tuple(x.shape), x.abs().max().item()))
    torch.save(dict(weight_seed=11, cases=cases), os.path.join(GOLD, "codec_decoder.pt"))
    print("  %-28s %8.1f KB" % ("codec_decoder.pt", os.path.getsize(os.path.join(GOLD, "codec_decoder.pt")) / 1024))


def make_codec_encoder_golden():
    """Encodec-48k ENCODER + residual vector quantizer: Hugging Face port (EncodecEncoder, EncodecResidualVectorQuantizer),
    seeded weights / codebooks regenerated from the seed by jen1_b200.codec_config.random_encoder_state_dict."""
    from transformers import EncodecConfig
    from transformers.models.encodec.modeling_encodec import EncodecEncoder, EncodecResidualVectorQuantizer
    from jen1_b200.codec_config import CodecDesc, random_encoder_state_dict
    desc = CodecDesc()
    cfg = EncodecConfig(sampling_rate=48000, audio_channels=desc.channels, normalize=True, chunk_length_s=1.0, overlap=0.01,
                        hidden_size=desc.dimension, num_filters=desc.n_filters, num_residual_layers=1,
                        upsampling_ratios=list(desc.ratios), norm_type="time_group_norm", kernel_size=desc.kernel_size,
                        last_kernel_size=desc.last_kernel_size, residual_kernel_size=desc.residual_kernel_size,
                        dilation_growth_rate=2, use_causal_conv=False, pad_mode="reflect", compress=desc.compress,
                        num_lstm_layers=desc.lstm_layers, trim_right_ratio=1.0, use_conv_shortcut=True,
                        target_bandwidths=[3.0, 6.0, 12.0, 24.0], codebook_size=1024)
    enc = EncodecEncoder(cfg).eval()
    rvq = EncodecResidualVectorQuantizer(cfg).eval()
    assert rvq.num_quantizers == 16
    sd = random_encoder_state_dict(desc, 21)
    hf = {}
    for k, v in sd.items():
        if k.startswith("encoder."):
            n = k[len("encoder."):].replace("model.", "layers.", 1).replace(".conv.conv.", ".conv.").replace(".conv.norm.", ".norm.")
            hf[n] = v
    enc.load_state_dict(hf, strict=True)
    for i in range(16):
        rvq.layers[i].codebook.embed.copy_(sd["quantizer.vq.layers.%d._codebook.embed" % i])
    cases = {}
    for name, N, L, seed in (("n2_l6400", 2, 6400, 5), ("n1_l5000", 1, 5000, 6), ("n3_l963", 3, 963, 7)):
        g = torch.Generator().manual_seed(seed)
        a = torch.randn(N, desc.channels, L, generator=g) * 0.5
        with torch.no_grad():
            emb = enc(a)
            codes = rvq.encode(emb)          # [n_q, N, T]
            q = rvq.decode(codes)            # [N, D, T]
        cases[name] = dict(audio=a, emb=emb.clone(), codes=codes.clone().to(torch.int32), quantized=q.clone())
        print("  encoder %-9s emb %s codes %s |emb| max %.3f" % (name, tuple(emb.shape), tuple(codes.shape), emb.abs().max().item()))
    torch.save(dict(weight_seed=21, cases=cases), os.path.join(GOLD, "codec_encoder.pt"))
    print("  %-28s %8.1f KB" % ("codec_encoder.pt", os.path.getsize(os.path.join(GOLD, "codec_encoder.pt")) / 1024))


def main():
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "codec":
        make_codec_golden()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "codec_encoder":
        make_codec_encoder_golden()
        return
    ref_import.install_shims()
    if len(sys.argv) > 1 and sys.argv[1] == "c2":
        make_config2_golden()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "cond":
        make_conditioner_golden()
        return

    # ---- 1. state_dict inventory of the reference (names, shapes, order) -------------------------------
    spec = {}
    for tag, desc in (("full", UNetDesc()), ("tiny", tiny_desc())):
        model = ref_import.build_reference_unet(**desc_overrides(desc))
        spec[tag] = [[k, list(v.shape)] for k, v in model.state_dict().items()]
        assert [(n, list(s)) for n, s, _ in desc.tensor_spec()] == [(k, s) for k, s in spec[tag]], tag
    with open(os.path.join(GOLD, "state_dict_spec.json"), "w") as f:
        json.dump(spec, f)

    # ---- 2. tiny UNet: several shapes / variants + per-stage taps --------------------------------------
    desc = tiny_desc()
    sd = random_state_dict(desc, seed=7)
    model = ref_import.build_reference_unet(**desc_overrides(desc))
    model.load_state_dict(sd, strict=True)
    cases = {
        "T50_B2": (2, 50, 11, 0, ["plain", "cfg", "cfg_causal", "cfg_noscale"]),
        "T33_B1_masked": (1, 33, 12, 5, ["plain", "cfg"]),
        "T8_B3": (3, 8, 13, 0, ["cfg", "cfg_causal"]),
        "T1_B1": (1, 1, 14, 3, ["cfg"]),
    }
    tiny = run_unet_cases(desc, model, cases)
    torch.save(dict(weights_seed=7, cases=tiny), os.path.join(GOLD, "unet_tiny.pt"))

    # ---- 3. diffusion process on the tiny model --------------------------------------------------------
    diff = ref_import.build_reference_diffusion(sampling_steps=100)
    tables = {k: getattr(diff, k).clone() for k in (
        "betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
        "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
        "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2")}
    pairs = {}
    for S in (100, 50, 25, 250, 999):
        times = torch.linspace(-1, 999, steps=S + 1)
        times = list(reversed(times.int().tolist()))
        pairs[str(S)] = list(zip(times[:-1], times[1:]))
    gdm = dict(tables=tables, pairs=pairs)
    from jen1.diffusion.gdm.noise_schedule import get_beta_schedule
    gdm["betas_cosine"] = get_beta_schedule("cosine", 1000)[0].to(torch.float32)

    # DDIM trajectories with the reference loop driving the reference tiny model
    traj = {}
    for tag, (B, T, S, seed, causal, use_init) in {
        "S25_B2_T50": (2, 50, 25, 21, False, False),
        "S25_B1_T33_causal_init": (1, 33, 25, 22, True, True),
    }.items():
        d = ref_import.build_reference_diffusion(sampling_steps=S)
        x, t, emb, mask, cc = make_inputs(desc, B, T, seed, 4)
        cond = dict(cross_attn_cond=emb, cross_attn_masks=mask, global_cond=None, input_concat_cond=cc)
        init = 0.5 * x if use_init else None
        torch.manual_seed(seed)
        y = d.sample(model, (B, desc.in_channels, T), cond, return_all_timesteps=True, causal=causal, init_data=init)
        traj[tag] = dict(B=B, T=T, S=S, seed=seed, causal=causal, use_init=use_init, all_steps=y.clone())
    gdm["traj"] = traj

    # q_sample / training loss
    d = ref_import.build_reference_diffusion(sampling_steps=100)
    x, t, emb, mask, cc = make_inputs(desc, 3, 50, 31, 0)
    cond = dict(cross_attn_cond=emb, cross_attn_masks=mask, global_cond=None, input_concat_cond=cc)
    noise = torch.randn_like(x)
    gdm["q_sample"] = dict(seed=31, x_t=d.q_sample(x, t, noise).clone(), noise=noise)
    torch.manual_seed(77)
    with torch.no_grad():
        gdm["train_loss"] = dict(seed=31, rng_seed=77, loss=float(d.training_loosses(model, x, t, cond, causal=False)))
    for obj in ("x0", "v"):
        d2 = ref_import.build_reference_diffusion(sampling_steps=25, objective=obj)
        xx, tt, emb, mask, cc = make_inputs(desc, 1, 20, 41, 0)
        cond = dict(cross_attn_cond=emb, cross_attn_masks=mask, global_cond=None, input_concat_cond=cc)
        torch.manual_seed(5)
        gdm["traj_" + obj] = d2.sample(model, (1, desc.in_channels, 20), cond).clone()
    torch.save(gdm, os.path.join(GOLD, "gdm.pt"))

    # ---- 4. full-size UNet at config 1 (T=150) ---------------------------------------------------------
    desc = UNetDesc()
    sd = random_state_dict(desc, seed=0)
    model = ref_import.build_reference_unet()
    model.load_state_dict(sd, strict=True)
    cases = {
        "T150_B1": (1, 150, 101, 0, ["plain", "cfg", "cfg_causal"]),
        "T150_B1_masked": (1, 150, 102, 100, ["cfg"]),
        "T47_B2": (2, 47, 103, 0, ["cfg"]),
    }
    full = run_unet_cases(desc, model, cases)
    for rec in full.values():  # keep the fixture small: fp32 [B,128,T]
        pass
    torch.save(dict(weights_seed=0, cases=full), os.path.join(GOLD, "unet_full.pt"))
    make_config2_golden()
    make_conditioner_golden()
    print("golden fixtures written to", GOLD)
    for fn in sorted(os.listdir(GOLD)):
        print("  %-28s %8.1f KB" % (fn, os.path.getsize(os.path.join(GOLD, fn)) / 1024))


if __name__ == "__main__":
    main()
