"""`Jen1` inference facade with the reference's signature (generation.py:17-132).

    Jen1(ckpt_path, device, sample_rate, cross_attn_cond_ids, global_cond_ids, input_concat_ids)
    .generate(prompt, seed=-1, steps=100, batch_size=1, seconds=30, use_gdm=False, task='text_guided',
              init_audio=None, init_audio_sr=None, inpainting_scope=None) -> Tensor

The hot loop `diffusion.sample(model, shape, conditioning, ...)` runs on the B200 engine.  The reference's
glue around it has several defects; this facade implements the INTENDED semantics and
each deviation is listed in DESIGN.md:
  * the model/engine is built once, not on every generate() call (generation.py:62-72);
  * `causal` reaches the sampler as a keyword (the reference passes it into `return_all_timesteps`);
  * `use_gdm=False` selects the reference's VDM sampler, which is non-functional; GDM/DDIM is always used;
  * per-sample masks / latents keep their batch dimension (generation.py:173-180 drops it);
  * Encodec (pip `encodec`, not installed, weights unreachable) is an injectable `codec` object with
    `encode_latent(audio)->[B,128,T]`, `decode_latent(latent)->[B,2,samples]`; `codec_state_dict=` (the Encodec model's
    state_dict, or only its decoder's) builds the B200 codec engines (jen1_b200/codec.py: decoder, and encoder + residual
    vector quantizer when the encoder's tensors are present); without either the facade works in the latent domain
    (`init_latent=` in, latents out).
"""
from __future__ import annotations

import math
import warnings
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .conditioners import MultiConditioner, RandomTextConditioner
from .config import DiffusionDesc, UNetDesc, latent_frames
from .diffusion import GaussianDiffusion, get_beta_schedule
from .model import UNetCFG1d
from .weights import load_checkpoint_state_dict, random_state_dict


def resample_frac(x: torch.Tensor, old_sr: int, new_sr: int, zeros: int = 24, rolloff: float = 0.945) -> torch.Tensor:
    """Fractional resampling by windowed-sinc polyphase filtering -- the published algorithm of `julius.resample_frac`
    (julius 0.2.x, the dependency `encodec.utils.convert_audio` calls; reference generation.py:95): reduce old_sr / new_sr
    by their gcd, low-pass at rolloff * min(old, new) / 2 with `zeros` zero crossings and a Hann^2-shaped (cos^2) window,
    one FIR phase per output sample of a period, replicate padding, output length floor(new_sr * L / old_sr).
    julius is not installed in this image: parity with it is UNPINNED (checked only through resampling identities)."""
    old_sr, new_sr = int(old_sr), int(new_sr)
    if old_sr == new_sr:
        return x
    g = math.gcd(old_sr, new_sr)
    old_sr, new_sr = old_sr // g, new_sr // g
    sr = min(new_sr, old_sr) * rolloff
    width = math.ceil(zeros * old_sr / sr)
    idx = torch.arange(-width, width + old_sr, dtype=torch.float32)
    kernels = []
    for i in range(new_sr):
        t = ((-i / new_sr + idx / old_sr) * sr).clamp_(-zeros, zeros) * math.pi
        window = torch.cos(t / zeros / 2) ** 2
        k = torch.sinc(t / math.pi) * window
        kernels.append(k / k.sum())
    kernel = torch.stack(kernels).view(new_sr, 1, -1).to(x)
    shape = x.shape
    length = shape[-1]
    y = torch.nn.functional.pad(x.reshape(-1, 1, length), (width, width + old_sr), mode="replicate")
    y = torch.nn.functional.conv1d(y, kernel, stride=old_sr)            # [*, new_sr, time]
    y = y.transpose(1, 2).reshape(*shape[:-1], -1)
    return y[..., : int(new_sr * length / old_sr)]


def convert_audio(wav: torch.Tensor, sr: int, target_sr: int, target_channels: int) -> torch.Tensor:
    """`encodec.utils.convert_audio` (reference generation.py:95): channel adaptation (mono <- mean, mono -> expand,
    equal -> as is), then resampling to the model rate."""
    assert wav.dim() >= 2 and wav.shape[-2] in (1, 2), "audio must be [..., channels (1 or 2), samples]"
    *lead, channels, length = wav.shape
    if target_channels == 1:
        wav = wav.mean(-2, keepdim=True)
    elif target_channels == 2:
        wav = wav.expand(*lead, target_channels, length)
    elif channels == 1:
        wav = wav.expand(target_channels, -1)
    else:
        raise RuntimeError("Impossible to convert from %d to %d channels" % (channels, target_channels))
    return resample_frac(wav, sr, target_sr)


class Jen1:
    def __init__(self, ckpt_path: Optional[str], device="cuda:0", sample_rate: int = 48000,
                 cross_attn_cond_ids: Sequence[str] = ("prompt",), global_cond_ids: Sequence[str] = (),
                 input_concat_ids: Sequence[str] = ("masked_input", "mask"), *, desc: Optional[UNetDesc] = None,
                 diffusion: Optional[DiffusionDesc] = None, conditioner: Optional[MultiConditioner] = None,
                 codec=None, codec_state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 state_dict: Optional[Dict[str, torch.Tensor]] = None, dtype: str = "bf16",
                 random_init_seed: Optional[int] = None, rng_device=None, use_cuda_graph: bool = True):
        self.ckpt_path, self.device, self.sample_rate = ckpt_path, torch.device(device), sample_rate
        self.cross_attn_cond_ids = list(cross_attn_cond_ids)
        self.global_cond_ids = list(global_cond_ids)
        self.input_concat_ids = list(input_concat_ids)
        assert not self.global_cond_ids, "global conditioning is not part of the reference configuration"
        self.desc = desc or UNetDesc()
        self.dcfg = diffusion or DiffusionDesc()
        self.conditioner = conditioner or MultiConditioner(
            {"prompt": RandomTextConditioner(self.desc.context_embedding_features, self.desc.context_embedding_max_length)})
        if codec is None and codec_state_dict is not None:
            from .codec import EncodecCodec
            codec = EncodecCodec(codec_state_dict, device=self.device)
        self.codec = codec
        self.rng_device, self.use_cuda_graph = rng_device, use_cuda_graph
        if state_dict is None:
            if ckpt_path is not None:
                state_dict = load_checkpoint_state_dict(ckpt_path, self.desc)
            elif random_init_seed is not None:
                state_dict = random_state_dict(self.desc, random_init_seed)
            else:
                raise ValueError("Jen1 needs ckpt_path, state_dict or random_init_seed")
        self.model = UNetCFG1d(self.desc, device=self.device, dtype=dtype).load_state_dict(state_dict)
        self._diffusions: Dict[int, GaussianDiffusion] = {}

    # reference generation.py:36-74 (GDM branch); cached per step count
    def get_model_and_diffusion(self, steps: int, use_gdm: bool = True):
        if steps not in self._diffusions:
            c = self.dcfg
            betas, alphas = get_beta_schedule(c.noise_schedule, c.steps)
            self._diffusions[steps] = GaussianDiffusion(
                steps=c.steps, betas=betas.to(torch.float32), alphas=alphas, objective=c.objective,
                loss_type=c.loss_type, device=self.device, cfg_dropout_proba=c.cfg_dropout_proba,
                embedding_scale=c.embedding_scale, batch_cfg=c.batch_cfg, scale_cfg=c.scale_cfg,
                sampling_timesteps=steps, ddim_sampling_eta=c.ddim_sampling_eta, use_fp16=False,
                scale_phi=c.scale_phi, rng_device=self.rng_device, use_cuda_graph=self.use_cuda_graph)
        return self._diffusions[steps], self.model

    def get_mask(self, sample_size: int, start: float, end: float, batch_size: int) -> torch.Tensor:
        """reference generation.py:134-143: ones with zeros over [start, end) seconds; 0 = region to generate."""
        mask = torch.ones((batch_size, 1, sample_size))
        mask[:, :, math.floor(start * self.sample_rate): math.ceil(end * self.sample_rate)] = 0
        return mask

    def get_emb(self, audio: torch.Tensor) -> torch.Tensor:
        if self.codec is None:
            raise RuntimeError("encoding audio needs a codec (Encodec is not bundled); pass init_latent= instead")
        return self.codec.encode_latent(audio)

    def get_conditioning(self, cond: Dict) -> Dict:
        """reference generation.py:152-192 with the batch dimension kept."""
        ca = torch.cat([cond[k][0] for k in self.cross_attn_cond_ids], dim=1) if self.cross_attn_cond_ids else None
        cm = torch.cat([cond[k][1] for k in self.cross_attn_cond_ids], dim=1) if self.cross_attn_cond_ids else None
        ic = torch.cat([cond[k] for k in self.input_concat_ids], dim=1) if self.input_concat_ids else None
        return {"cross_attn_cond": ca, "cross_attn_masks": cm, "global_cond": None, "input_concat_cond": ic}

    @torch.no_grad()
    def generate(self, prompt, seed: int = -1, steps: int = 100, batch_size: int = 1, seconds: float = 30,
                 use_gdm: bool = False, task: str = "text_guided", init_audio: Optional[torch.Tensor] = None,
                 init_audio_sr: Optional[int] = None, inpainting_scope=None, *,
                 init_latent: Optional[torch.Tensor] = None, return_latents: Optional[bool] = None):
        if not use_gdm:
            warnings.warn("the reference's VDM sampler is non-functional (SURVEY.md 3.6); using GDM/DDIM", stacklevel=2)
        seed = seed if seed != -1 else int(np.random.randint(0, 2 ** 32 - 1))
        torch.manual_seed(seed)
        diffusion, model = self.get_model_and_diffusion(steps, True)
        B = batch_size
        T = latent_frames(seconds, self.sample_rate)
        C = self.desc.in_channels
        dev = self.device
        prompts = list(prompt) if isinstance(prompt, (list, tuple)) else [prompt] * B
        assert len(prompts) == B

        if init_audio is not None and init_latent is None:
            if init_audio.dim() == 2:
                init_audio = init_audio.unsqueeze(0).repeat(B, 1, 1)
            channels = getattr(self.codec, "channels", init_audio.shape[1])
            init_audio = convert_audio(init_audio, init_audio_sr or self.sample_rate, self.sample_rate, channels)
            init_latent = self.get_emb(init_audio.to(dev))
        if init_latent is not None:
            init_latent = init_latent.to(dev, torch.float32)
            if init_latent.dim() == 2:
                init_latent = init_latent.unsqueeze(0).repeat(B, 1, 1)

        sample_length = int(round(seconds * self.sample_rate))
        init_data = None
        if task == "text_guided":
            mask = self.get_mask(sample_length, 0, seconds, B)
            causal = False
            latent = torch.zeros(B, C, T, device=dev)
        elif task == "music_inpaint":
            assert init_latent is not None and inpainting_scope is not None
            assert init_latent.shape[-1] == T, "init latent must cover `seconds`"
            mask = self.get_mask(sample_length, inpainting_scope[0], inpainting_scope[1], B)
            causal = False
            latent = init_data = init_latent
        elif task == "music_cont":
            assert init_latent is not None
            Tc = init_latent.shape[-1]
            assert Tc < T, "continuation needs a prefix shorter than `seconds`"
            start_s = Tc / T * seconds
            mask = self.get_mask(sample_length, start_s, seconds, B)
            causal = True
            latent = torch.cat([init_latent, torch.zeros(B, C, T - Tc, device=dev)], dim=2)
            init_data = latent
        else:
            raise ValueError(f"unknown task {task}")
        mask = torch.nn.functional.interpolate(mask.to(dev), size=T)  # nearest, reference generation.py:117
        cond = self.conditioner([{"prompt": p} for p in prompts], dev)
        cond["masked_input"] = latent * mask
        cond["mask"] = mask
        conditioning = self.get_conditioning(cond)
        latents = diffusion.sample(model, (B, C, T), conditioning, causal=causal, init_data=init_data)
        if return_latents or (return_latents is None and self.codec is None):
            return latents
        return self.codec.decode_latent(latents)
