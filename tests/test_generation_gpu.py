"""`Jen1.generate()` facade on the GPU (reference generation.py:76-132, intended semantics -- DESIGN.md section 8):
the three tasks run through the engine, `causal` reaches the sampler for continuation, masks are resampled to the
latent rate per sample, and a seed reproduces the output.  Latent-domain unless a codec is attached (last test: the sampled
latent goes through the Encodec-decoder engine).
"""
import pytest
import torch

from jen1_b200.config import latent_frames, tiny_desc
from jen1_b200.weights import random_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def jen():
    from jen1_b200.generation import Jen1
    desc = tiny_desc()
    return Jen1(None, device=DEV, desc=desc, state_dict=random_state_dict(desc, 7), dtype="fp32")


def test_text_guided_shapes_and_seed(jen):
    secs, B = 0.5, 2
    T = latent_frames(secs)
    a = jen.generate(["a calm piano", "drums"], seed=3, steps=20, batch_size=B, seconds=secs, use_gdm=True)
    b = jen.generate(["a calm piano", "drums"], seed=3, steps=20, batch_size=B, seconds=secs, use_gdm=True)
    c = jen.generate(["a calm piano", "drums"], seed=4, steps=20, batch_size=B, seconds=secs, use_gdm=True)
    assert a.shape == (B, jen.desc.in_channels, T) and torch.isfinite(a).all()
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert a.abs().max().item() <= 1.0 + 1e-6  # the last DDIM step returns the clamped x0


def test_inpaint_equals_manual_conditioning(jen):
    """generate(task='music_inpaint') == diffusion.sample() on the conditioning the reference INTENDS to build:
    input_concat_cond = cat(latent * mask, mask), init_data = latent, non-causal."""
    secs, B, steps = 0.5, 2, 20
    T, C = latent_frames(secs), jen.desc.in_channels
    g = torch.Generator().manual_seed(9)
    lat = (torch.randn(B, C, T, generator=g) * 0.5).to(DEV)
    out = jen.generate(["x", "y"], seed=11, steps=steps, batch_size=B, seconds=secs, use_gdm=True,
                       task="music_inpaint", init_latent=lat, inpainting_scope=(0.125, 0.375))
    mask = jen.get_mask(int(round(secs * 48000)), 0.125, 0.375, B)
    mask = torch.nn.functional.interpolate(mask.to(DEV), size=T)
    assert 0 < mask.sum().item() < mask.numel()
    torch.manual_seed(11)
    cond = jen.conditioner([{"prompt": p} for p in ["x", "y"]], DEV)
    cond["masked_input"], cond["mask"] = lat * mask, mask
    dif, model = jen.get_model_and_diffusion(steps, True)
    ref = dif.sample(model, (B, C, T), jen.get_conditioning(cond), causal=False, init_data=lat)
    assert torch.equal(out, ref)


def test_continuation_is_causal(jen):
    secs, B, steps = 0.5, 1, 20
    T, C = latent_frames(secs), jen.desc.in_channels
    g = torch.Generator().manual_seed(2)
    prefix = (torch.randn(B, C, T // 2, generator=g) * 0.5).to(DEV)
    out = jen.generate("z", seed=5, steps=steps, batch_size=B, seconds=secs, use_gdm=True, task="music_cont",
                       init_latent=prefix)
    assert out.shape == (B, C, T) and torch.isfinite(out).all()
    # same call through the sampler with causal=False must differ: the flag reaches the engine
    full = torch.cat([prefix, torch.zeros(B, C, T - T // 2, device=DEV)], 2)
    mask = torch.nn.functional.interpolate(jen.get_mask(int(round(secs * 48000)), (T // 2) / T * secs, secs, B).to(DEV), size=T)
    torch.manual_seed(5)
    cond = jen.conditioner([{"prompt": "z"}], DEV)
    cond["masked_input"], cond["mask"] = full * mask, mask
    dif, model = jen.get_model_and_diffusion(steps, True)
    noncausal = dif.sample(model, (B, C, T), jen.get_conditioning(cond), causal=False, init_data=full)
    torch.manual_seed(5)
    cond = jen.conditioner([{"prompt": "z"}], DEV)
    cond["masked_input"], cond["mask"] = full * mask, mask
    causal = dif.sample(model, (B, C, T), jen.get_conditioning(cond), causal=True, init_data=full)
    assert torch.equal(out, causal) and not torch.equal(out, noncausal)


def test_generate_decodes_through_the_codec_engine():
    """With a codec attached, generate() returns audio: the sampled latent goes through the B200 Encodec-decoder engine
    (reference generation.py:128-131), and equals decoding the latent-domain result of the same seed."""
    from jen1_b200.codec import EncodecCodec
    from jen1_b200.codec_config import CodecDesc, random_state_dict as codec_sd
    from jen1_b200.generation import Jen1
    desc = tiny_desc()
    cdesc = CodecDesc(channels=2, dimension=desc.in_channels, n_filters=4, ratios=(4, 2))
    codec = EncodecCodec(codec_sd(cdesc, 5), cdesc, DEV)
    jen = Jen1(None, device=DEV, desc=desc, state_dict=random_state_dict(desc, 7), dtype="fp32", codec=codec)
    secs, B = 0.5, 2
    T = latent_frames(secs)
    audio = jen.generate(["a", "b"], seed=3, steps=10, batch_size=B, seconds=secs, use_gdm=True)
    lat = jen.generate(["a", "b"], seed=3, steps=10, batch_size=B, seconds=secs, use_gdm=True, return_latents=True)
    assert audio.shape == (B, 2, T * cdesc.hop) and torch.isfinite(audio).all()
    assert torch.equal(audio, codec.decode_latent(lat))


def test_continuation_from_audio_goes_through_the_encoder_engine():
    """generate(task='music_cont', init_audio=...) with a full Encodec state_dict: prompt audio -> Encodec encoder + residual
    vector quantizer engine (reference generation.py:95,145-150) -> causal sampling -> decoder engine -> audio."""
    from jen1_b200.codec_config import CodecDesc, random_encoder_state_dict, random_state_dict as codec_sd
    from jen1_b200.config import UNetDesc
    from jen1_b200.generation import Jen1
    desc = UNetDesc(in_channels=128, channels=32, multipliers=(1, 1, 2, 2, 4), factors=(1, 4, 2, 2), num_blocks=(1, 2, 2, 1),
                    attentions=(0, 0, 1, 1), out_channels=128, context_channels=(129,), context_embedding_features=64,
                    context_embedding_max_length=12, attention_heads=4)
    cdesc = CodecDesc()
    csd = dict(codec_sd(cdesc, 11))
    csd.update(random_encoder_state_dict(cdesc, 21))
    jen = Jen1(None, device=DEV, desc=desc, state_dict=random_state_dict(desc, 3), dtype="fp32", codec_state_dict=csd)
    assert jen.codec.encoder is not None
    prompt_audio = torch.randn(2, 48000, generator=torch.Generator().manual_seed(1)) * 0.2  # 1 s of stereo
    n0 = jen.codec.encoder.launch_count()
    audio = jen.generate("x", seed=5, steps=5, batch_size=1, seconds=2, use_gdm=True, task="music_cont",
                         init_audio=prompt_audio, init_audio_sr=48000)
    assert jen.codec.encoder.launch_count() > n0
    T = latent_frames(2)
    assert audio.shape == (1, 2, T * 320) and torch.isfinite(audio).all()
    lat = jen.codec.encode_latent(prompt_audio.unsqueeze(0))
    assert lat.shape == (1, 128, latent_frames(1))  # 150 frames + 2 of the 1 % overlap segment (reference segment loop)
