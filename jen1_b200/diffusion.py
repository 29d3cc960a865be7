"""Gaussian diffusion process with the reference's call surface (jen1/diffusion/gdm/gdm.py:14-272).

`GaussianDiffusion(steps=..., betas=..., objective=..., ...)` takes the reference constructor arguments and
`sample(model, shape, conditioning, return_all_timesteps=False, causal=False, init_data=None)` is the
reference seam (gdm.py:228-230): `model` may be ANY callable with the reference model signature.  When it is
a `jen1_b200.model.UNetCFG1d` and DDIM sampling is selected, the loop body (UNet evaluation, classifier-free
guidance, x0/eps conversion, clamp, DDIM update) runs as one CUDA graph per step inside the engine; otherwise
the generic loop below drives the callable with torch ops (A/B harness, e.g. the reference model itself).

Bit-exact pieces (host side, computed with the same torch fp32 ops as the reference): the schedule tables
(gdm.py:54-87, noise_schedule.py:7-30), the DDIM index list (gdm.py:190-193) and the per-step scalars
alpha/sigma/c (gdm.py:212-216).  All random draws are made with torch in the reference's order
(randn(shape); per step bernoulli (model.py:325) then randn_like (gdm.py:218)); `rng_device="cpu"` makes them
on the CPU generator so a CUDA run reproduces a CPU reference run with the same seed.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .model import UNetCFG1d


def get_beta_schedule(schedule_name: str, num_diffusion_timesteps: int):
    """reference jen1/diffusion/gdm/noise_schedule.py:7-30; returns (betas, None) like the reference."""
    if schedule_name == "linear":
        scale = 1000 / num_diffusion_timesteps
        return torch.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps), None
    if schedule_name == "cosine":
        def alpha_bar(t):
            return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        n = num_diffusion_timesteps
        return torch.tensor([min(1 - alpha_bar((i + 1) / n) / alpha_bar(i / n), 0.999) for i in range(n)]), None
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def extract(a: torch.Tensor, t: torch.Tensor, x_shape) -> torch.Tensor:
    """reference utils/script_util.py:43-46."""
    out = a.gather(-1, t)
    return out.reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))


class GaussianDiffusion:
    def __init__(self, *, steps, betas, objective="noise", loss_type="l2", device="cuda", cfg_dropout_proba=0.1,
                 embedding_scale=0.8, batch_cfg=False, scale_cfg=False, sampling_timesteps=None,
                 ddim_sampling_eta=1.0, use_fp16=False, alphas=None, scale_phi=0.7, rng_device=None,
                 use_cuda_graph=True):
        assert objective in {"noise", "x0", "v"}
        assert loss_type in {"l1", "l2"}
        self.objective, self.loss_type = objective, loss_type
        self.device = torch.device(device)
        self.cfg_dropout_proba, self.embedding_scale = cfg_dropout_proba, embedding_scale
        self.batch_cfg, self.scale_cfg, self.scale_phi = batch_cfg, scale_cfg, scale_phi
        self.use_fp16 = use_fp16  # kept for signature parity; the engine's precision is set on the model
        self.rng_device = rng_device
        self.use_cuda_graph = use_cuda_graph
        self.shard = None  # (lo, hi, global_batch) while a sharded sample() is running (jen1_b200/sharding.py)
        self.num_timesteps = steps
        self.sampling_timesteps = steps if sampling_timesteps is None else sampling_timesteps
        assert self.sampling_timesteps <= self.num_timesteps
        self.is_ddim_sampling = self.sampling_timesteps < self.num_timesteps
        self.ddim_sampling_eta = ddim_sampling_eta
        # schedule tables on the host in fp32, same op sequence as the reference (gdm.py:54-87)
        betas = betas.detach().to("cpu", torch.float32)
        assert betas.dim() == 1 and (betas > 0).all() and (betas <= 1).all()
        alphas = (1 - betas) if alphas is None else alphas.detach().to("cpu", torch.float32)
        self.betas = betas
        self.alphas_cumprod = torch.cumprod(alphas, dim=0)
        self.alphas_cumprod_prev = F.pad(self.alphas_cumprod[:-1], (1, 0), value=1.0)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = torch.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = torch.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = torch.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = torch.log(
            torch.cat([self.posterior_variance[1].unsqueeze(0), self.posterior_variance[1:]]))
        self.posterior_mean_coef1 = betas * torch.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * torch.sqrt(alphas) / (1.0 - self.alphas_cumprod)

    # nn.Module-like no-ops used by the reference glue
    def eval(self):
        return self

    def to(self, device):
        self.device = torch.device(device)
        return self

    # ---- integer index arithmetic (bit-exact) ------------------------------------------------------------
    def time_pairs(self) -> List[Tuple[int, int]]:
        """reference gdm.py:190-193."""
        times = torch.linspace(-1, self.num_timesteps - 1, steps=self.sampling_timesteps + 1)
        times = list(reversed(times.int().tolist()))
        return list(zip(times[:-1], times[1:]))

    def ddim_coefficients(self) -> torch.Tensor:
        """Per-step scalars [S, 8] (see include/jen1_b200.h), computed as reference gdm.py:212-216 does."""
        rows = []
        for time, time_next in self.time_pairs():
            row = [self.sqrt_recip_alphas_cumprod[time], self.sqrt_recipm1_alphas_cumprod[time],
                   self.sqrt_alphas_cumprod[time], self.sqrt_one_minus_alphas_cumprod[time]]
            if time_next < 0:
                row += [torch.tensor(0.0)] * 3 + [torch.tensor(1.0)]
            else:
                alpha, alpha_next = self.alphas_cumprod[time], self.alphas_cumprod[time_next]
                sigma = self.ddim_sampling_eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
                c = (1 - alpha_next - sigma ** 2).sqrt()
                row += [alpha_next.sqrt(), c, sigma, torch.tensor(0.0)]
            rows.append(torch.stack([r.to(torch.float32) for r in row]))
        return torch.stack(rows)

    # ---- x0 / eps / v conversions (reference gdm.py:89-105) ----------------------------------------------
    def _tab(self, name: str, like: torch.Tensor) -> torch.Tensor:
        return getattr(self, name).to(like.device)

    def predict_start_from_noise(self, x_t, t, noise):
        return (extract(self._tab("sqrt_recip_alphas_cumprod", x_t), t, x_t.shape) * x_t
                - extract(self._tab("sqrt_recipm1_alphas_cumprod", x_t), t, x_t.shape) * noise)

    def predict_noise_from_start(self, x_t, t, x0):
        return ((extract(self._tab("sqrt_recip_alphas_cumprod", x_t), t, x_t.shape) * x_t - x0)
                / extract(self._tab("sqrt_recipm1_alphas_cumprod", x_t), t, x_t.shape))

    def predict_start_from_v(self, x_t, t, v):
        return (extract(self._tab("sqrt_alphas_cumprod", x_t), t, x_t.shape) * x_t
                - extract(self._tab("sqrt_one_minus_alphas_cumprod", x_t), t, x_t.shape) * v)

    def q_posterior(self, x_start, x_t, t):
        mean = (extract(self._tab("posterior_mean_coef1", x_t), t, x_t.shape) * x_start
                + extract(self._tab("posterior_mean_coef2", x_t), t, x_t.shape) * x_t)
        return (mean, extract(self._tab("posterior_variance", x_t), t, x_t.shape),
                extract(self._tab("posterior_log_variance_clipped", x_t), t, x_t.shape))

    def _call_model(self, model, x, t, conditioning, causal):
        extra = {}
        if isinstance(model, UNetCFG1d) and self.cfg_dropout_proba > 0.0:
            # the cond-dropout bernoulli (reference model.py:325) is drawn HERE, with this object's rng_device /
            # shard rules, so the seed-parity guarantees also hold when the model is driven step by step
            extra["drop_mask"] = self._bernoulli(x.shape[0], x.device)
        return model(x, t, **extra, embedding=conditioning["cross_attn_cond"], embedding_mask=conditioning["cross_attn_masks"],
                     embedding_scale=self.embedding_scale, embedding_mask_proba=self.cfg_dropout_proba,
                     features=conditioning["global_cond"], channels_list=[conditioning["input_concat_cond"]],
                     batch_cfg=self.batch_cfg, scale_cfg=self.scale_cfg, causal=causal)

    def model_predictions(self, x, t, model, conditioning=None, clip_x_start=False, causal=False):
        """reference gdm.py:116-142."""
        model_out = self._call_model(model, x, t, conditioning, causal)
        clip = (lambda v: torch.clamp(v, min=-1, max=1.0)) if clip_x_start else (lambda v: v)
        if self.objective == "noise":
            pred_noise = model_out
            x_start = clip(self.predict_start_from_noise(x, t, pred_noise))
        elif self.objective == "x0":
            x_start = clip(model_out)
            pred_noise = self.predict_noise_from_start(x, t, x_start)
        else:
            x_start = clip(self.predict_start_from_v(x, t, model_out))
            pred_noise = self.predict_noise_from_start(x, t, x_start)
        return pred_noise, x_start

    # ---- RNG helpers ---------------------------------------------------------------------------------------
    def _draw_shape(self, shape):
        """With `self.shard = (lo, hi, B)` (jen1_b200.sharding) the FULL-batch tensor is drawn and sliced, so a
        sharded run consumes the same random stream as the single-device run."""
        sh = self.shard
        if sh is not None and len(shape) > 0 and shape[0] == sh[1] - sh[0]:
            return (sh[2],) + tuple(shape[1:]), slice(sh[0], sh[1])
        return tuple(shape), None

    def _randn(self, shape, device):
        full, sl = self._draw_shape(shape)
        if self.rng_device is not None and torch.device(self.rng_device) != torch.device(device):
            out = torch.randn(full, device=self.rng_device)
            return (out if sl is None else out[sl]).to(device)
        out = torch.randn(full, device=device)
        return out if sl is None else out[sl].contiguous()

    def _bernoulli(self, b, device):
        p = float(self.cfg_dropout_proba)
        if p == 1:
            return torch.ones((b, 1, 1), device=device, dtype=torch.bool)
        rd = device if self.rng_device is None else self.rng_device
        full, sl = self._draw_shape((b, 1, 1))
        out = torch.bernoulli(torch.full(full, p, device=rd)).to(torch.bool)
        return (out if sl is None else out[sl]).to(device)

    # ---- samplers ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def ddim_sample(self, model, shape, conditioning, return_all_timesteps=False, causal=False, init_data=None):
        """reference gdm.py:181-225."""
        if isinstance(model, UNetCFG1d):
            return self._ddim_sample_engine(model, shape, conditioning, return_all_timesteps, causal, init_data)
        device = self.device
        audio = self._randn(shape, device)
        if init_data is not None:
            audio = audio + init_data.to(device)
        audios = [audio]
        ac = self.alphas_cumprod
        for time, time_next in self.time_pairs():
            time_cond = torch.full((shape[0],), time, device=device, dtype=torch.long)
            pred_noise, x_start = self.model_predictions(audio, time_cond, model, conditioning, clip_x_start=True,
                                                         causal=causal)
            audios.append(audio)
            if time_next < 0:
                audio = x_start
                continue
            alpha, alpha_next = ac[time], ac[time_next]
            sigma = self.ddim_sampling_eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
            c = (1 - alpha_next - sigma ** 2).sqrt()
            noise = self._randn(tuple(audio.shape), device)
            audio = x_start * alpha_next.sqrt().to(device) + c.to(device) * pred_noise + sigma.to(device) * noise
        return audio if not return_all_timesteps else torch.stack(audios, dim=1)

    @torch.no_grad()
    def _ddim_sample_engine(self, model: UNetCFG1d, shape, conditioning, return_all_timesteps, causal, init_data):
        """The same loop with the body executed by the engine (one CUDA graph replay per step)."""
        eng = model.engine
        if eng is None:
            raise RuntimeError("UNetCFG1d has no weights loaded")
        assert conditioning.get("global_cond") is None, "global conditioning is not supported"
        device = eng.device
        B, Cc, T = shape
        pairs = self.time_pairs()
        # always rebuild the prompt K/V cache at the start of a trajectory (one small GEMM against 100 steps): the
        # model-side cache key cannot see `.data` swaps or a direct engine.set_context with another batch
        model.set_context(conditioning["cross_attn_cond"], conditioning["cross_attn_masks"], force=True)
        eng.set_timesteps([t for t, _ in pairs])
        use_cfg = self.embedding_scale != 1.0
        # CUDA graphs cannot be captured on the legacy default stream: run the loop on a side stream
        cur = torch.cuda.current_stream(device)
        side = model.__dict__.get("_side_stream")
        if side is None or side.device != device:
            side = model.__dict__["_side_stream"] = torch.cuda.Stream(device)  # one capture / replay stream per model
        side.wait_stream(cur)
        with torch.cuda.device(device), torch.cuda.stream(side):
            eng.sample_begin(self.ddim_coefficients(), conditioning["input_concat_cond"], B, T, causal,
                             float(self.embedding_scale) if use_cfg else 1.0, self.scale_cfg, self.scale_phi,
                             self.objective, self.use_cuda_graph)
            audio = self._randn(tuple(shape), device)
            if init_data is not None:
                audio = audio + init_data.to(device)
            # persistent state / noise buffers per shape: the engine's captured step graph points at them, so a
            # second sample() call of the same shape replays the graph instead of re-capturing it
            bufs = model.__dict__.setdefault("_sampler_buffers", {})
            key = (tuple(shape), str(device))
            if key not in bufs:
                bufs.clear()
                bufs[key] = (torch.empty(tuple(shape), device=device, dtype=torch.float32),
                             torch.empty(tuple(shape), device=device, dtype=torch.float32))
            x, noise = bufs[key]
            x.copy_(audio)
            audios = [x.clone()] if return_all_timesteps else None
            for i, (time, time_next) in enumerate(pairs):
                drop = None
                if self.cfg_dropout_proba > 0.0:
                    drop = self._bernoulli(B, device).reshape(B).contiguous()
                if return_all_timesteps:
                    audios.append(x.clone())
                last = time_next < 0
                if not last:
                    if self.shard is not None or (self.rng_device is not None and torch.device(self.rng_device) != device):
                        noise.copy_(self._randn(tuple(x.shape), device))
                    else:
                        noise.normal_()
                eng.sample_step(i, x, noise, drop)
            out = x.clone() if not return_all_timesteps else torch.stack(audios, dim=1)
        cur.wait_stream(side)
        out.record_stream(cur)
        return out

    @torch.no_grad()
    def p_sample_loop(self, model, shape, conditioning, return_all_timesteps=False, causal=False, init_data=None):
        """Ancestral sampler (reference gdm.py:153-179).  The reference version raises TypeError when reached
        through `sample()`; this is the intended loop, including the reference's
        UNIFORM posterior noise (`torch.rand_like`, gdm.py:161)."""
        device = self.device
        audio = self._randn(shape, device)
        if init_data is not None:
            audio = audio + init_data.to(device)
        audios = [audio]
        for t in reversed(range(0, self.num_timesteps)):
            bt = torch.full((shape[0],), t, device=device, dtype=torch.long)
            _, x_start = self.model_predictions(audio, bt, model, conditioning, clip_x_start=False, causal=causal)
            x_start = x_start.clamp(-1.0, 1.0)
            mean, _, logvar = self.q_posterior(x_start, audio, bt)
            noise = torch.rand_like(audio) if t > 0 else 0.0
            audio = mean + (0.5 * logvar).exp() * noise
            audios.append(audio)
        return audio if not return_all_timesteps else torch.stack(audios, dim=1)

    @torch.no_grad()
    def sample(self, model, shape, conditioning, return_all_timesteps=False, causal=False, init_data=None):
        """reference gdm.py:227-230."""
        fn = self.ddim_sample if self.is_ddim_sampling else self.p_sample_loop
        return fn(model, shape, conditioning, return_all_timesteps=return_all_timesteps, causal=causal,
                  init_data=init_data)

    # ---- training-side pieces (reference gdm.py:232-272) ---------------------------------------------------
    def q_sample(self, x_start, t, noise=None):
        if noise is None:
            noise = torch.rand_like(x_start)  # uniform, as in the reference (gdm.py:237)
        assert noise.shape == x_start.shape
        return (extract(self._tab("sqrt_alphas_cumprod", x_start), t, x_start.shape) * x_start
                + extract(self._tab("sqrt_one_minus_alphas_cumprod", x_start), t, x_start.shape) * noise)

    def training_losses(self, model, x_start, t, conditioning, noise=None, causal=False):
        if noise is None:
            noise = torch.rand_like(x_start)
        x_t = self.q_sample(x_start, t, noise=noise)
        model_out = self._call_model(model, x_t, t, conditioning, causal)
        if self.objective == "noise":
            target = noise
        elif self.objective == "x0":
            target = x_start
        else:
            target = (extract(self._tab("sqrt_alphas_cumprod", x_start), t, x_start.shape) * noise
                      - extract(self._tab("sqrt_one_minus_alphas_cumprod", x_start), t, x_start.shape) * x_start)
        fn = F.l1_loss if self.loss_type == "l1" else F.mse_loss
        loss = fn(model_out, target, reduction="none")
        return loss.reshape(loss.shape[0], -1).mean(dim=1).mean()

    training_loosses = training_losses  # the reference's spelling (gdm.py:245)


def create_gaussian_diffusion(steps=1000, noise_schedule="linear", objective="v", loss_type="l2", device="cuda",
                              cfg_dropout_proba=0.1, embedding_scale=1, batch_cfg=False, scale_cfg=False,
                              sampling_steps=None, use_fp16=False, **extra) -> GaussianDiffusion:
    """reference utils/script_util.py:216-249."""
    betas, alphas = get_beta_schedule(noise_schedule, steps)
    return GaussianDiffusion(steps=steps, betas=betas.to(torch.float32), alphas=alphas, objective=objective,
                             loss_type=loss_type, device=device, cfg_dropout_proba=cfg_dropout_proba,
                             embedding_scale=embedding_scale, batch_cfg=batch_cfg, scale_cfg=scale_cfg,
                             sampling_timesteps=sampling_steps, use_fp16=use_fp16, **extra)
