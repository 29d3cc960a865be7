"""Host-side logic of the PRODUCT package against the reference goldens, on CPU (no engine involved):
schedule tables and DDIM index arithmetic bit-exact, per-step DDIM scalars, q_sample / training loss, and the generic
sampling loop (`GaussianDiffusion.sample` driving an arbitrary callable -- here the oracle denoiser) against the
reference trajectories."""
import os

import pytest
import torch

from jen1_b200.config import UNetDesc, latent_frames, tiny_desc
from jen1_b200.diffusion import create_gaussian_diffusion, extract, get_beta_schedule
from jen1_b200.weights import random_state_dict
from oracle.make_golden import make_inputs
from oracle.unet_oracle import OracleUNet


def _dif(S=100, objective="noise", **kw):
    return create_gaussian_diffusion(steps=1000, noise_schedule="linear", objective=objective, device="cpu",
                                     cfg_dropout_proba=0.2, embedding_scale=0.8, batch_cfg=True, scale_cfg=True,
                                     sampling_steps=S, **kw)


def test_product_schedule_tables_bit_exact(golden_dir):
    fx = torch.load(os.path.join(golden_dir, "gdm.pt"))
    d = _dif(100)
    for k, ref in fx["tables"].items():
        assert torch.equal(getattr(d, k), ref), k
    assert torch.equal(get_beta_schedule("cosine", 1000)[0].to(torch.float32), fx["betas_cosine"])
    for S, pairs in fx["pairs"].items():
        assert _dif(int(S)).time_pairs() == [tuple(p) for p in pairs]


def test_ddim_coefficients_follow_reference_formulas():
    d = _dif(100)
    coef = d.ddim_coefficients()
    pairs = d.time_pairs()
    assert coef.shape == (100, 8) and pairs[0][0] == 999 and pairs[-1] == (9, -1)
    ac = d.alphas_cumprod
    for i in (0, 1, 50, 98):
        t, tn = pairs[i]
        a, an = ac[t], ac[tn]
        sigma = 1.0 * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()  # reference gdm.py:214
        c = (1 - an - sigma ** 2).sqrt()
        assert torch.equal(coef[i, 0], d.sqrt_recip_alphas_cumprod[t]) and torch.equal(coef[i, 1], d.sqrt_recipm1_alphas_cumprod[t])
        assert torch.equal(coef[i, 4], an.sqrt()) and torch.equal(coef[i, 5], c) and torch.equal(coef[i, 6], sigma)
        assert coef[i, 7] == 0
    assert coef[99, 7] == 1 and torch.all(coef[99, 4:7] == 0)  # last step returns x0 (gdm.py:208-210)


def test_extract_gathers_per_sample():
    a = torch.arange(10.0)
    out = extract(a, torch.tensor([3, 7]), (2, 4, 5))
    assert out.shape == (2, 1, 1) and out.flatten().tolist() == [3.0, 7.0]


def test_product_q_sample_and_loss_match_reference(golden_dir):
    fx = torch.load(os.path.join(golden_dir, "gdm.pt"))
    desc = tiny_desc()
    model = OracleUNet(desc, random_state_dict(desc, 7))
    d = _dif(100)
    x, t, emb, mask, cc = make_inputs(desc, 3, 50, 31, 0)
    assert torch.equal(d.q_sample(x, t, fx["q_sample"]["noise"]), fx["q_sample"]["x_t"])
    cond = dict(cross_attn_cond=emb, cross_attn_masks=mask, global_cond=None, input_concat_cond=cc)
    torch.manual_seed(fx["train_loss"]["rng_seed"])
    loss = float(d.training_loosses(model, x, t, cond))
    assert abs(loss - fx["train_loss"]["loss"]) < 1e-5


def test_product_generic_loop_matches_reference_trajectories(golden_dir):
    fx = torch.load(os.path.join(golden_dir, "gdm.pt"))
    desc = tiny_desc()
    model = OracleUNet(desc, random_state_dict(desc, 7))
    tag, rec = next(iter(fx["traj"].items()))
    x, t, emb, mask, cc = make_inputs(desc, rec["B"], rec["T"], rec["seed"], 4)
    cond = dict(cross_attn_cond=emb, cross_attn_masks=mask, global_cond=None, input_concat_cond=cc)
    init = 0.5 * x if rec["use_init"] else None
    torch.manual_seed(rec["seed"])
    y = _dif(rec["S"]).sample(model, (rec["B"], desc.in_channels, rec["T"]), cond, return_all_timesteps=True,
                              causal=rec["causal"], init_data=init)
    assert y.shape == rec["all_steps"].shape
    assert (y - rec["all_steps"]).abs().max().item() < 5e-4, tag
    for obj in ("x0", "v"):
        xx, tt, emb, mask, cc = make_inputs(desc, 1, 20, 41, 0)
        cond = dict(cross_attn_cond=emb, cross_attn_masks=mask, global_cond=None, input_concat_cond=cc)
        torch.manual_seed(5)
        y = _dif(25, objective=obj).sample(model, (1, desc.in_channels, 20), cond)
        assert (y - fx["traj_" + obj]).abs().max().item() < 5e-4, obj


def test_latent_frame_algebra_and_level_lengths():
    assert latent_frames(10) == 1515 and latent_frames(30) == 4545
    d = UNetDesc()
    assert d.level_lengths(4545) == [4545, 4545, 1137, 285, 72, 36, 18, 9, 5, 3]
    assert d.level_lengths(1515) == [1515, 1515, 379, 95, 24, 12, 6, 3, 2, 1]
    assert d.num_params() == 296543106


def test_int_and_number_conditioners_match_reference_golden(golden_dir):
    """reference jen1/conditioners.py:114-164 (+ utils/module.py NumberEmbedder) on the reference's own seeded weights."""
    import os

    import torch

    from jen1_b200.conditioners import IntConditioner, NumberConditioner
    fx = torch.load(os.path.join(golden_dir, "conditioners.pt"))
    ic = IntConditioner(64, 0, 512)
    ic.int_embedder.weight.data.copy_(fx["int_sd"]["int_embedder.weight"])
    nc = NumberConditioner(64, 0, 512)
    nc.embedder.load_reference_state_dict(fx["num_sd"])
    with torch.no_grad():
        io, no = ic(fx["ints"], "cpu"), nc(fx["floats"], "cpu")
    assert torch.equal(io[0], fx["int_out"][0]) and torch.equal(io[1], fx["int_out"][1])
    assert torch.allclose(no[0], fx["num_out"][0], atol=1e-6, rtol=0) and torch.equal(no[1], fx["num_out"][1])


def test_multi_conditioner_factory_builds_every_type():
    """reference utils/script_util.py:151-178, intended semantics (the reference returns inside its loop)."""
    from jen1_b200.conditioners import RandomTextConditioner, create_multi_conditioner
    mc = create_multi_conditioner(text_conditioner=RandomTextConditioner())
    out = mc([{"prompt": "a song", "seconds_start": 3, "seconds_total": 100.0},
              {"prompt": ["wrapped"], "seconds_start": 600, "seconds_total": 30}], "cpu")
    assert set(out) == {"prompt", "seconds_start", "seconds_total"}
    assert out["prompt"][0].shape == (2, 128, 1024) and out["prompt"][1].shape == (2, 128)
    assert out["seconds_start"][0].shape == (2, 1, 1024) and out["seconds_total"][0].shape == (2, 1, 1024)
    import pytest
    with pytest.raises(ValueError, match="not found in batch metadata"):
        mc([{"prompt": "x"}], "cpu")


def test_convert_audio_resampling_identities():
    """`convert_audio` (reference generation.py:95 -> encodec.utils / julius.resample_frac; parity unpinned, julius is not
    installed): output length floor(new * L / old), DC preserved, a band-limited sine survives 44.1k -> 48k -> 44.1k, and
    channel adaptation follows the encodec rules."""
    import math

    import torch

    from jen1_b200.generation import convert_audio, resample_frac
    L, old, new = 44100, 44100, 48000
    t = torch.arange(L) / old
    x = torch.stack([torch.sin(2 * math.pi * 440 * t), torch.ones(L) * 0.3]).unsqueeze(0)  # [1, 2, L]
    y = resample_frac(x, old, new)
    assert y.shape == (1, 2, int(new * L / old))
    assert (y[0, 1, 100:-100] - 0.3).abs().max() < 1e-4
    t2 = torch.arange(y.shape[-1]) / new
    assert (y[0, 0, 500:-500] - torch.sin(2 * math.pi * 440 * t2)[500:-500]).abs().max() < 2e-3
    back = resample_frac(y, new, old)
    n = min(back.shape[-1], L)
    assert (back[0, 0, 500:n - 500] - x[0, 0, 500:n - 500]).abs().max() < 4e-3
    assert resample_frac(x, 48000, 48000) is x
    mono = convert_audio(x, 48000, 48000, 1)
    assert mono.shape == (1, 1, L) and torch.allclose(mono[0, 0], x[0].mean(0))
    st = convert_audio(x[:, :1], 48000, 48000, 2)
    assert st.shape == (1, 2, L) and torch.equal(st[0, 0], st[0, 1])


def test_codec_description_key_layouts_and_work_model():
    """Encodec decoder host logic: the tensor inventory of the 48 kHz decoder, both state_dict key layouts (pip package /
    Hugging Face port, optional `decoder.` prefix), shape checking, and the work model bench.py reports against."""
    from jen1_b200.codec_config import CodecDesc, canonical_state_dict, decode_work, random_state_dict as codec_sd, to_hf_names
    desc = CodecDesc()
    assert desc.hidden == 512 and desc.hop == 320
    kinds = [k for _, k, *_ in desc.layers()]
    assert kinds == ["conv", "lstm"] + ["convtr", "res"] * 4 + ["conv"]
    assert [i for i, *_ in desc.layers()] == [0, 1, 3, 4, 6, 7, 9, 10, 12, 13, 15]  # model.N indices (ELUs hold no tensors)
    sd = codec_sd(desc, 1)
    assert sum(v.numel() for v in sd.values()) == sum(int(torch.tensor(s).prod()) for _, s, _ in desc.tensor_spec())
    assert sd["model.3.convtr.convtr.weight"].shape == (512, 256, 16) and sd["model.1.lstm.weight_hh_l1"].shape == (2048, 512)
    hf = {"decoder." + k: v for k, v in to_hf_names(sd).items()}
    assert "decoder.layers.4.block.1.conv.weight" in hf and "decoder.layers.4.shortcut.norm.bias" in hf
    back = canonical_state_dict(desc, dict(hf, **{"encoder.layers.0.conv.weight": torch.zeros(1)}))  # extra keys ignored
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    bad = dict(sd)
    bad["model.0.conv.conv.weight"] = torch.zeros(512, 128, 5)
    with pytest.raises(ValueError):
        canonical_state_dict(desc, bad)
    w = decode_work(desc, 4545)  # 181 GFLOP per 30 s sample
    assert abs(w["flops"] / 1e9 - 181.2) < 0.5 and w["samples"] == 4545 * 320
    assert 4.0e9 < w["bytes"] < 4.6e9
