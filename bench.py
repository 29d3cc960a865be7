#!/usr/bin/env python
"""Benchmark of the JEN-1 denoiser hot path (BASELINE.json metric: denoiser latent-frames/sec/step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload config3|config2|config5] [--scaling weak|strong]

One "step" = one DDIM sampler step of the hot path over one batch: pack x -> UNetCFG1d on the doubled CFG batch
-> CFG combine + std rescale -> x0/eps conversion + clamp -> DDIM update (reference gdm.py:202-222), plus the
per-step RNG draws the reference makes (bernoulli cond-dropout, randn_like).  value = B*T latent frames per GPU
* N GPUs / seconds-per-step.

Default workload = BASELINE.json configs[2] ("config3": 30 s latent T=4545, 32 prompts over 8 GPUs = 4 samples per
GPU) -- the configuration the north-star's targets are quoted on; it fits one GPU, so every N runs it.
`--scaling weak` (default) keeps 4 samples per GPU; `--scaling strong` keeps the global batch at 32 (32/N per GPU).

  * `value`  : K steps timed with CUDA events on the launching stream, everything resident in HBM.
  * `e2e`    : the same metric through the public API -- `GaussianDiffusion.sample(model, shape, conditioning)`
               with the conditioning in pinned HOST memory (H2D inside the timed region) and the finished latent
               read back to the host (D2H inside), K sampler steps per call; median of 7 calls.
  * `roofline`: whole-step HBM roofline (one CUDA-graph launch = one step): algorithmic bytes (jen1_b200.workload)
               / event-timed step duration vs MEASURED_PEAKS.json.
  * `cpu_baseline`: the reference's CPU path on this box's host cores (baseline/_ref when installed, else the oracle
               port), same workload, a bounded number of sampler steps; `parity` = rel-L2 between the engine and the
               CPU oracle for one CFG evaluation of the benchmark's own inputs.
  * `gpu_eager_baseline`: the oracle's functional PyTorch forward with its tensors on the same GPU (cuDNN / cuBLAS
               eager, then the same step under torch.cuda.graphs, fp32 with PyTorch's default TF32 flags and under bf16
               autocast) -- the library bar on this box, timed outside the product arm's timed region.
  * `--impl reference`: the reference CPU arm alone (rank 0 only): the UNMODIFIED reference from baseline/_ref
               (scripts/install_reference.py) driven through its own `GaussianDiffusion.sample`; the oracle port only
               if that directory is absent.

Multi-GPU (torchrun, one rank per GPU): independent samples sharded over ranks, NO data-path collective; the only
communication is the barrier + max-over-ranks of the timing.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: 100-step DDIM, 10 s @48 kHz Encodec latent, batch=1, 1xB200
    "config2": dict(B=1, seconds=10, T=1515, name="100-step DDIM, 10 s (T=1515) latent, batch %d per GPU, CFG 0.8 (%d UNet rows)"),
    # BASELINE.json configs[2]: 30 s, batch 32 over 8 GPUs = 4 samples per GPU
    "config3": dict(B=4, seconds=30, T=4545, name="100-step DDIM, 30 s (T=4545) latent, batch %d per GPU, CFG 0.8 (%d UNet rows)"),
    # BASELINE.json configs[4]: continuation (masked latent, causal convs + causal self-attention), 30 s, batch 16 over 4 GPUs
    "config5": dict(B=4, seconds=30, T=4545, causal=True,
                    name="100-step DDIM continuation (causal), 30 s (T=4545) masked latent, batch %d per GPU, CFG 0.8 (%d UNet rows)"),
}
STRONG_GLOBAL_BATCH = 32
METRIC = "denoiser latent-frames/sec/step"
UNIT = "latent-frames/s"
DATA = "synthetic (seeded random-init weights, N(0,1) latents/embeddings)"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tensor=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor=1590.0, source="fallback (B200_PROFILING.md)")


def _config(wl, world, scaling):
    """The `config` object -- identical in both arms (ours / reference) for the same command line."""
    from jen1_b200.config import UNetDesc
    from jen1_b200.workload import step_bytes
    alg = step_bytes(UNetDesc(), wl["B"], wl["T"], cfg=True, elem_bytes=2)["total_bytes"]
    return {"workload": wl["name"], "frames_per_step_per_gpu": wl["B"] * wl["T"],
            "parallelism": "dp%d (batch sharding, no collective)" % world,
            "scaling": scaling if scaling == "weak" else "strong (global batch %d)" % STRONG_GLOBAL_BATCH,
            "l2": "per-step working set (%.0f MB weights + activations in bf16) exceeds the 126 MB L2" % (alg / 1e6)}


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region.  NVML is polled in-process every ~2 ms (a 20-step region
    of 60 ms still yields ~30 samples); falls back to `nvidia-smi -lms 20` (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thread = index, [], None, None
        self.sm, self.mx, self.reasons, self.stop_flag, self.how = [], None, set(), False, None

    def _nvml_loop(self, nv, h):
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(h))
                for n, b in bits.items():
                    if r & b:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                phys = int(vis.split(",")[self.index])
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.how = "nvml, 2 ms poll"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return self
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.how = "nvidia-smi -lms 20"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.how is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi / NVML unavailable"], "samples": 0}
        if self.proc is None:
            self.stop_flag = True
            self.thread.join(timeout=1)
        else:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for r in self.rows:
                if len(r) < 9:
                    continue
                try:
                    self.sm.append(float(r[1]))
                    self.mx = float(r[2])
                except ValueError:
                    continue
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                "samples": len(sm), "how": self.how}


def _profile_facts(workload):
    """DRAM traffic per step and kernel shares from the committed ncu launch list of the same command (profiles/)."""
    p = os.path.join(ROOT, "profiles", "step_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload, {})
        except ValueError:
            pass
    return {}


def _make_problem(desc, B, T, seed, continuation=False):
    import torch
    g = torch.Generator().manual_seed(seed)
    emb = torch.randn(B, desc.context_embedding_max_length, desc.context_embedding_features, generator=g)
    mask = torch.ones(B, desc.context_embedding_max_length, dtype=torch.bool)
    cc = torch.zeros(B, desc.context_channels[0], T)  # text-guided: zero masked latent + zero mask
    if continuation:  # keep the first half of a stand-in latent, generate the rest
        lat = torch.randn(B, desc.in_channels, T, generator=g) * 0.5
        keep = torch.zeros(B, 1, T)
        keep[:, :, : T // 2] = 1.0
        cc = torch.cat([lat * keep, keep], dim=1)
    return emb, mask, cc


def _ddim_scalars(dif, time_, time_next):
    """gdm.py:212-216 as python floats (for the legs that drive a model callable step by step)."""
    a, an = dif.alphas_cumprod[time_], dif.alphas_cumprod[max(time_next, 0)]
    sigma = dif.eta * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
    c = (1 - an - sigma ** 2).sqrt()
    return (float(dif.sqrt_recip_alphas_cumprod[time_]), float(dif.sqrt_recipm1_alphas_cumprod[time_]), float(an.sqrt()),
            float(c), float(sigma))


def _callable_step(model_fn, dif, x, t, cond, scal, causal):
    """One DDIM step (noise objective) around any callable with the reference model signature."""
    import torch
    sr, srm1, san, c, sigma = scal
    eps = dif.model_call(model_fn, x, t, cond, causal)
    x0 = torch.clamp(sr * x - srm1 * eps, min=-1.0, max=1.0)
    return x0 * san + c * eps + sigma * torch.randn_like(x)


# ------------------------------------------------------------------------------------------------------------
def _reference_stack(sd, device="cpu"):
    """(model, make_diffusion(sampling_steps), kind): the UNMODIFIED reference from baseline/_ref when installed, else
    the oracle port."""
    from jen1_b200.config import UNetDesc
    from oracle import ref_import
    root = ref_import.reference_root(prefer_installed=True)
    if root is not None:
        import warnings
        warnings.filterwarnings("ignore")
        ref_import.install_shims(root)
        model = ref_import.build_reference_unet()
        model.load_state_dict(sd, strict=True)
        model = model.to(device).eval()
        return model, (lambda S: ref_import.build_reference_diffusion(sampling_steps=S, device=device)), "reference", root
    from oracle.gdm_oracle import OracleDiffusion
    from oracle.unet_oracle import OracleUNet
    return OracleUNet(UNetDesc(), sd), (lambda S: OracleDiffusion(sampling_timesteps=S)), "port", None


def run_reference(args, wl, scaling):
    """Reference CPU arm (rank 0 only; other ranks exit without work): `GaussianDiffusion.sample` of the reference on
    the host cores, W sampling steps untimed, then K sampling steps timed (bounded: K is cut so the arm ends within
    minutes; the line says how many steps were timed)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from jen1_b200.config import UNetDesc
    from jen1_b200.weights import random_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    desc = UNetDesc()
    B, T = wl["B"], wl["T"]
    sd = random_state_dict(desc, 0)
    model, make_dif, kind, root = _reference_stack(sd)
    causal = bool(wl.get("causal", False))
    emb, mask, cc = _make_problem(desc, B, T, 1, causal)
    cond = dict(cross_attn_cond=emb, cross_attn_masks=mask, global_cond=None, input_concat_cond=cc)
    shape = (B, desc.in_channels, T)
    torch.manual_seed(0)
    with torch.no_grad():
        t0 = time.perf_counter()
        make_dif(max(args.warmup, 1)).sample(model, shape, cond, causal=causal)
        warm_s = (time.perf_counter() - t0) / max(args.warmup, 1)
        budget = 150.0  # seconds of timed CPU work
        k_timed = max(3, min(args.steps, int(budget / max(warm_s, 1e-3))))
        t0 = time.perf_counter()
        make_dif(k_timed).sample(model, shape, cond, causal=causal)
        dt = time.perf_counter() - t0
    ms = dt / k_timed * 1e3
    val = B * T / (ms / 1e3)
    sample = ("%d DDIM sampling steps (of --steps %d; bounded to ~%d s of CPU work) of the full per-GPU workload (B=%d, T=%d, "
              "CFG -> %d UNet rows), fp32, through %s" % (k_timed, args.steps, int(budget), B, T, 2 * B,
              "the UNMODIFIED reference's GaussianDiffusion.sample (baseline/_ref)" if kind == "reference"
              else "the oracle port's DDIM loop (baseline/_ref not installed)"))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32", "data": DATA, "config": _config(wl, args.gpus, scaling),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "steps_timed": k_timed}
    print(json.dumps(line), flush=True)


def cpu_baseline(desc, sd, wl, budget_s=20.0):
    """The reference CPU path beside the GPU number (rank 0, N=1): warm median of bounded single steps."""
    import torch
    from oracle.gdm_oracle import OracleDiffusion
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, T = wl["B"], wl["T"]
    model, _, kind, _ = _reference_stack(sd)
    dif = OracleDiffusion(sampling_timesteps=100)
    causal = bool(wl.get("causal", False))
    emb, mask, cc = _make_problem(desc, B, T, 1, causal)
    cond = dict(cross_attn_cond=emb, cross_attn_masks=mask, global_cond=None, input_concat_cond=cc)
    pairs = dif.time_pairs()
    torch.manual_seed(0)
    x = torch.randn(B, desc.in_channels, T)
    times = []
    with torch.no_grad():
        t_start = time.perf_counter()
        i = 0
        while True:
            time_, time_next = pairs[i % (len(pairs) - 1)]
            tc = torch.full((B,), time_, dtype=torch.long)
            t0 = time.perf_counter()
            x = _callable_step(model, dif, x, tc, cond, _ddim_scalars(dif, time_, time_next), causal)
            times.append(time.perf_counter() - t0)
            i += 1
            if i >= 3 and (time.perf_counter() - t_start > budget_s or i >= 40):
                break
    warm = sorted(times[2:]) if len(times) > 3 else sorted(times)
    med = warm[len(warm) // 2]
    return {"value": B * T / med, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d sampler steps (median of the last %d) of the full workload B=%d T=%d, fp32, %s"
                      % (len(times), len(warm), B, T, "UNMODIFIED reference model (baseline/_ref) stepped by the DDIM update"
                         if kind == "reference" else "oracle port"), "ms_per_step": med * 1e3}


def parity_check(desc, sd, model, wl, dev):
    """rel-L2 between the engine (the arm being timed) and the CPU oracle for ONE CFG evaluation of this benchmark's
    inputs (no cond-dropout, so both sides evaluate the same thing)."""
    import torch
    from oracle.unet_oracle import unet_cfg_forward
    B, T = wl["B"], wl["T"]
    causal = bool(wl.get("causal", False))
    emb, mask, cc = _make_problem(desc, B, T, 1, causal)
    g = torch.Generator().manual_seed(77)
    x = torch.randn(B, desc.in_channels, T, generator=g)
    t = torch.full((B,), 989, dtype=torch.long)
    kw = dict(embedding_scale=0.8, batch_cfg=True, scale_cfg=True, embedding_mask_proba=0.0, causal=causal)
    with torch.no_grad():
        ref = unet_cfg_forward(desc, sd, x, t, embedding=emb, embedding_mask=mask, channels_list=[cc], **kw)
    y = model(x.to(dev), t.to(dev), embedding=emb.to(dev), embedding_mask=mask.to(dev), features=None,
              channels_list=[cc.to(dev)], **kw).cpu()
    return {"rel_l2": ((y - ref).norm() / ref.norm()).item(), "gate": 1e-2,
            "what": "engine (%s) vs CPU oracle fp32, one CFG evaluation of the benchmark inputs at t=989" % model.dtype}


def gpu_eager_baseline(desc, sd, wl, dev, steps=10):
    """SURVEY.md 8(c) last row: the library bar on the same box.  The oracle's functional PyTorch forward with weights
    and inputs on the GPU (cuDNN convs, cuBLAS GEMMs, ATen norms), one DDIM step per iteration: eager, then captured in
    a torch.cuda.graphs CUDA graph; fp32 (PyTorch defaults: cudnn.allow_tf32=True, matmul.allow_tf32=False) and bf16
    autocast.  Timed with CUDA events after the product arm has finished."""
    import torch
    from oracle.gdm_oracle import OracleDiffusion
    from oracle.unet_oracle import OracleUNet
    B, T = wl["B"], wl["T"]
    causal = bool(wl.get("causal", False))
    out = {"unit": "ms_per_step", "steps": steps,
           "flags": {"cudnn.allow_tf32": bool(torch.backends.cudnn.allow_tf32),
                     "cuda.matmul.allow_tf32": bool(torch.backends.cuda.matmul.allow_tf32)},
           "what": "oracle functional forward (PyTorch ops) with tensors on cuda, same workload and DDIM step"}
    try:
        sd_d = {k: v.to(dev) for k, v in sd.items()}
        model = OracleUNet(desc, sd_d)
        dif = OracleDiffusion(sampling_timesteps=100)
        emb, mask, cc = _make_problem(desc, B, T, 1, causal)
        cond = dict(cross_attn_cond=emb.to(dev), cross_attn_masks=mask.to(dev), global_cond=None, input_concat_cond=cc.to(dev))
        pairs = dif.time_pairs()
        scal = _ddim_scalars(dif, *pairs[1])
        tc = torch.full((B,), pairs[1][0], dtype=torch.long, device=dev)
        x = torch.randn(B, desc.in_channels, T, device=dev)

        def timed(fn, n):
            for _ in range(3):
                fn()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1) / n

        for tag, ctx in (("fp32", None), ("bf16_autocast", torch.bfloat16)):
            state = {"x": x.clone()}

            def step():
                with torch.no_grad():
                    if ctx is None:
                        state["x"].copy_(_callable_step(model, dif, state["x"], tc, cond, scal, causal))
                    else:
                        with torch.autocast("cuda", dtype=ctx):
                            state["x"].copy_(_callable_step(model, dif, state["x"], tc, cond, scal, causal).float())

            out[tag + "_eager_ms"] = timed(step, steps)
            try:
                s = torch.cuda.Stream(dev)
                s.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(s):
                    for _ in range(2):
                        step()
                torch.cuda.current_stream(dev).wait_stream(s)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    step()
                out[tag + "_cudagraph_ms"] = timed(g.replay, steps)
                del g
            except Exception as e:  # capture can fail on library ops that sync; the eager number stands
                out[tag + "_cudagraph_ms"] = None
                out[tag + "_cudagraph_error"] = str(e).splitlines()[0][:160]
        del model, sd_d
        torch.cuda.empty_cache()
    except Exception as e:
        out["error"] = str(e).splitlines()[0][:200]
    return out


# ------------------------------------------------------------------------------------------------------------
def _device_timed(model, dif, desc, wl, dev, rank, world, K, W, clock_index=None):
    """K sampler steps (CUDA graph per step) with inputs resident in HBM; returns (ms over K steps, launches, lps, clk)."""
    import torch
    import torch.distributed as dist
    eng = model.engine
    B, T = wl["B"], wl["T"]
    causal = bool(wl.get("causal", False))
    emb_h, mask_h, cc_h = _make_problem(desc, B, T, 1 + rank, causal)
    side = torch.cuda.Stream(dev)
    with torch.cuda.stream(side):
        emb, mask, cc = emb_h.to(dev), mask_h.to(dev), cc_h.to(dev)
        model.set_context(emb, mask, force=True)
        eng.set_timesteps([t for t, _ in dif.time_pairs()])
        coef = dif.ddim_coefficients()
        coef[:, 7] = 0.0  # no "last step" shortcut inside the timed window: every step does the full update
        coef[:, 4:7] = torch.nan_to_num(coef[:, 4:7])
        eng.sample_begin(coef, cc, B, T, causal, 0.8, True, 0.7, "noise", True)
        torch.manual_seed(1234 + rank)
        x = torch.randn(B, desc.in_channels, T, device=dev)
        noise = torch.empty_like(x)
        pfull = torch.full((B, 1, 1), 0.2, device=dev)

        def one_step(i):
            drop = torch.bernoulli(pfull).to(torch.bool).reshape(B)  # reference model.py:325
            noise.normal_()                                           # reference gdm.py:218
            eng.sample_step(i, x, noise, drop)

        l0 = eng.launch_count()
        for i in range(W):
            one_step(i)
        side.synchronize()
        lps = (eng.launch_count() - l0) // max(W, 1)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        clocks = ClockSampler(clock_index).start() if clock_index is not None else None
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l1 = eng.launch_count()
        ev0.record(side)
        for i in range(K):
            one_step(W + i)
        ev1.record(side)
        side.synchronize()
        torch.cuda.synchronize(dev)
        dev_ms = ev0.elapsed_time(ev1)
        launches = eng.launch_count() - l1
        clk = clocks.stop() if clocks else None
        if world > 1:
            dist.barrier()
    return dev_ms, launches, lps, clk


def codec_decode_leg(B, T, dev, cpu=True):
    """the Encodec-48k decoder engine on the finished latents of this workload (once per
    generate(), after the sampling loop -- NOT part of `value`), next to the CPU oracle on a bounded sample."""
    import torch
    from jen1_b200.codec import EncodecDecoder
    from jen1_b200.codec_config import CodecDesc, random_state_dict as codec_sd
    from jen1_b200.codec_config import decode_work
    cdesc = CodecDesc()
    csd = codec_sd(cdesc, 11)
    z = torch.randn(B, cdesc.dimension, T, generator=torch.Generator().manual_seed(3)).to(dev)

    def timed(precision, reps):
        dec = EncodecDecoder(cdesc, dev, precision).load_state_dict(csd)
        for _ in range(2):
            dec(z)
        torch.cuda.synchronize(dev)
        n0, t0, l0 = dec.launch_count(), dec.tf32_launch_count(), dec.lstm_tc_launch_count()
        each = []
        for _ in range(reps):  # one event pair per decode, median: a single host / driver hiccup must not move the number
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dec(z)
            e1.record()
            torch.cuda.synchronize(dev)
            each.append(e0.elapsed_time(e1))
        each.sort()
        return (dec, each[len(each) // 2], (dec.launch_count() - n0) // reps, (dec.tf32_launch_count() - t0) // reps,
                (dec.lstm_tc_launch_count() - l0) // reps, each)

    _, ms_strict, _, _, _, _ = timed("fp32", 3)
    dec, ms, nl, ntf, nlstm, each = timed("tf32", 7)
    work = decode_work(cdesc, T)
    pk = _peaks()
    leg = {"ms_per_decode": ms, "ms_per_decode_min_max": [each[0], each[-1]], "decodes_timed": len(each), "precision": "tf32 tensor-core convs, fp16 recurrent LSTM weights, fp32 storage / accumulation",
           "ms_per_decode_fp32_strict": ms_strict, "audio_seconds": B * T / 150.0,
           "samples_per_s": B * T * cdesc.hop / (ms / 1e3), "launches_per_decode": nl, "tf32_gemm_launches": ntf,
           "lstm_tensor_core_launches": nlstm, "lstm_cluster_ctas": dec.lstm_cluster(), "workspace_gb": dec.workspace_bytes(B, T) / 1e9,
           "algorithmic_gb": B * work["bytes"] / 1e9, "hbm_frac": B * work["bytes"] / (ms / 1e3) / 1e9 / pk["hbm"],
           "tflops": B * work["flops"] / (ms / 1e3) / 1e12,
           "what": "EncodecDecoder(latent [%d,128,%d]) -> audio [%d,2,%d] fp32, device-timed, latent resident; seeded "
                   "random-init weights (the pip checkpoint is unreachable offline)" % (B, T, B, T * cdesc.hop)}
    if cpu:
        from oracle.codec_oracle import decoder_forward
        Ts = max(32, T // 16)
        zc = z[:1, :, :Ts].cpu()
        with torch.no_grad():
            decoder_forward(cdesc, csd, zc[:, :, :16])
            t0 = time.perf_counter()
            ref = decoder_forward(cdesc, csd, zc)
            dt = time.perf_counter() - t0
        got = dec(zc.to(dev)).cpu()
        leg["cpu_oracle"] = {"ms_per_decode_scaled": dt * 1e3 * (B * T) / Ts, "cores": torch.get_num_threads(),
                             "sample": "oracle/codec_oracle.py on 1 x %d frames (%.1f s of CPU), scaled by frames" % (Ts, dt),
                             "parity_rel_l2": ((got - ref).norm() / ref.norm()).item()}
    return leg


def codec_encode_leg(B, T, dev):
    """The encode side of the codec (reference generation.py:145-150, used by the inpaint / continuation tasks): B clips of
    T/150 s -> quantised latents through EncodecCodec.encode_latent (segmentation + normalisation on the host side in torch,
    encoder + residual vector quantizer on the engine)."""
    import torch
    from jen1_b200.codec import EncodecCodec
    from jen1_b200.codec_config import CodecDesc, random_encoder_state_dict, random_state_dict as codec_sd
    cdesc = CodecDesc()
    sd = dict(codec_sd(cdesc, 11))
    sd.update(random_encoder_state_dict(cdesc, 21))
    codec = EncodecCodec(sd, cdesc, dev)
    seconds = T / 151.5
    L = int(round(seconds * 48000))
    audio = (torch.randn(B, 2, L, generator=torch.Generator().manual_seed(5)) * 0.3).to(dev)
    for _ in range(2):
        lat = codec.encode_latent(audio)
    torch.cuda.synchronize(dev)
    each = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lat = codec.encode_latent(audio)
        e1.record()
        torch.cuda.synchronize(dev)
        each.append(e0.elapsed_time(e1))
    each.sort()
    assert tuple(lat.shape) == (B, 128, T) and bool(torch.isfinite(lat).all())
    return {"ms_per_encode": each[len(each) // 2], "ms_per_encode_min_max": [each[0], each[-1]], "audio_seconds": B * seconds,
            "segments": B * len(range(0, L, codec.stride)),
            "what": "EncodecCodec.encode_latent(audio [%d,2,%d]) -> quantised latent [%d,128,%d]: 1 s segments (1 %% overlap, "
                    "loudness-normalised) batched through the encoder engine, 16-stage residual VQ; device-timed" % (B, L, B, T)}


def generate_audio_leg(desc, sd, B, T, steps, dev, dtype):
    """The reference's public entry point end to end (generation.py:76-132): Jen1.generate(prompts, steps, seconds) ->
    audio on the host.  Conditioner (seeded random text embeddings: T5 weights are unreachable offline), sampling loop on
    the UNet engine, Encodec decoder engine, device -> host copy of the audio; wall clock, median of 3 calls."""
    import torch
    from jen1_b200.codec_config import CodecDesc, random_state_dict as codec_sd
    from jen1_b200.generation import Jen1
    import warnings
    jen = Jen1(None, device=dev, desc=desc, state_dict=sd, dtype=dtype, codec_state_dict=codec_sd(CodecDesc(), 11))
    seconds = T / 151.5  # latent_frames(seconds) == T for the benchmark shapes (10 s -> 1515, 30 s -> 4545)
    prompts = ["prompt %d" % i for i in range(B)]
    times = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for rep in range(4):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            audio = jen.generate(prompts, seed=rep, steps=steps, batch_size=B, seconds=seconds, use_gdm=True).cpu()
            times.append((time.perf_counter() - t0) * 1e3)
    assert audio.shape == (B, 2, T * 320) and bool(torch.isfinite(audio).all())
    ms = sorted(times[1:])[1]
    return {"ms_per_call": ms, "audio_seconds": B * T / 150.0, "realtime_factor": B * T / 150.0 / (ms / 1e3), "steps": steps,
            "what": "Jen1.generate(%d prompts, steps=%d, seconds=%.0f) -> host audio [%d, 2, %d]: conditioner, DDIM loop on the "
                    "UNet engine, Encodec decoder engine, D2H copy; wall clock, median of 3 calls after one warm call"
                    % (B, steps, seconds, B, T * 320)}


def run_ours(args, wl, scaling):
    import torch
    import torch.distributed as dist
    from jen1_b200.config import UNetDesc
    from jen1_b200.diffusion import create_gaussian_diffusion
    from jen1_b200.model import UNetCFG1d
    from jen1_b200.weights import random_state_dict
    from jen1_b200.workload import step_bytes

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torchrun --nproc-per-node %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    desc = UNetDesc()
    B, T = wl["B"], wl["T"]
    K, W = args.steps, args.warmup
    causal = bool(wl.get("causal", False))
    sd = random_state_dict(desc, 0)
    model = UNetCFG1d(desc, device=dev, dtype=args.dtype).load_state_dict(sd)
    S = max(K + W + 1, 20)  # schedule length (the is-last row is never used): one conditioning row / coefficient row per step
    dif = create_gaussian_diffusion(steps=1000, noise_schedule="linear", objective="noise", device=dev,
                                    cfg_dropout_proba=0.2, embedding_scale=0.8, batch_cfg=True, scale_cfg=True,
                                    sampling_steps=S)

    # ---------------- device-resident timing: K sampler steps, CUDA events on the launching stream
    dev_ms, launches, launches_per_step, clk = _device_timed(model, dif, desc, wl, dev, rank, world, K, W,
                                                             clock_index=local if rank == 0 else None)
    t_ms = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = t_ms.item() / K
    value = world * B * T / (ms_per_step / 1e3)

    # ---------------- end to end through the public API with host buffers (median of 5 calls)
    emb_h, mask_h, cc_h = _make_problem(desc, B, T, 1 + rank, causal)
    emb_p, mask_p, cc_p = emb_h.pin_memory(), mask_h.pin_memory(), cc_h.pin_memory()
    dif_e = create_gaussian_diffusion(steps=1000, noise_schedule="linear", objective="noise", device=dev,
                                      cfg_dropout_proba=0.2, embedding_scale=0.8, batch_cfg=True, scale_cfg=True,
                                      sampling_steps=max(K, 20))
    Se = dif_e.sampling_timesteps
    out_h = torch.empty(B, desc.in_channels, T).pin_memory()

    def e2e_call():
        cond = dict(cross_attn_cond=emb_p.to(dev, non_blocking=True), cross_attn_masks=mask_p.to(dev, non_blocking=True),
                    global_cond=None, input_concat_cond=cc_p.to(dev, non_blocking=True))
        lat = dif_e.sample(model, (B, desc.in_channels, T), cond, causal=causal)
        out_h.copy_(lat, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    e2e = None
    if not args.no_e2e:
        import gc
        e2e_call()  # warm (graph capture for this shape, allocator)
        e2e_call()
        reps = []
        gc.collect()
        gc.disable()  # a generational collection in the middle of a call is host time the GPU queue cannot hide
        for _ in range(7):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            e2e_call()
            torch.cuda.synchronize(dev)
            t_e = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
            reps.append(t_e.item())
        gc.enable()
        e_ms_step = sorted(reps)[len(reps) // 2] / Se
        h2d = (emb_p.numel() * 4 + mask_p.numel() + cc_p.numel() * 4)
        d2h = out_h.numel() * 4
        e2e = {"value": world * B * T / (e_ms_step / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d / Se,
               "d2h_bytes_per_step": d2h / Se, "steps_per_call": Se, "ms_per_step": e_ms_step, "calls": len(reps),
               "ms_per_step_min_max": [min(reps) / Se, max(reps) / Se],
               "api": "GaussianDiffusion.sample(UNetCFG1d, shape, conditioning): pinned-host conditioning in, host latent out; "
                      "includes the per-call context K/V hoist and timestep tables; median of %d calls (max over ranks each)" % len(reps)}

    # ---------------- extras on a single GPU: config 2 riding along, the "smallest change" integration path
    extra = {}
    if world == 1 and not args.quick and args.workload != "config2" and scaling == "weak":
        wl2 = dict(WORKLOADS["config2"])
        ms2, _, lps2, _ = _device_timed(model, dif, desc, wl2, dev, 0, 1, K, W)
        extra["config2"] = {"workload": wl2["name"] % (wl2["B"], 2 * wl2["B"]), "ms_per_step": ms2 / K,
                            "value": wl2["B"] * wl2["T"] / (ms2 / K / 1e3), "unit": UNIT, "launches_per_step": int(lps2)}
    if world == 1 and not args.quick:
        # the reference's own loop calling model(x, t, ...) -> jen1_unet_forward every step
        emb_d, mask_d, cc_d = emb_h.to(dev), mask_h.to(dev), cc_h.to(dev)
        xx = torch.randn(B, desc.in_channels, T, device=dev)
        tt = torch.full((B,), 500, dtype=torch.long, device=dev)
        kw = dict(embedding=emb_d, embedding_mask=mask_d, features=None, channels_list=[cc_d], embedding_scale=0.8,
                  embedding_mask_proba=0.2, batch_cfg=True, scale_cfg=True, causal=causal)
        for _ in range(3):
            model(xx, tt, **kw)
        torch.cuda.synchronize(dev)
        n_f = max(10, min(K, 50))
        t0 = time.perf_counter()
        for _ in range(n_f):
            model(xx, tt, **kw)
        torch.cuda.synchronize(dev)
        f_ms = (time.perf_counter() - t0) * 1e3 / n_f
        extra["unet_forward_path"] = {"ms_per_call": f_ms, "value": B * T / (f_ms / 1e3), "unit": UNIT, "calls": n_f,
                                      "what": "UNetCFG1d.__call__ -> jen1_unet_forward per step (no sampler fusion), wall clock"}

    if world == 1 and not args.quick:
        extra["codec_decode"] = codec_decode_leg(B, T, dev, cpu=not args.no_cpu_baseline)
        extra["codec_encode"] = codec_encode_leg(B, T, dev)
        if args.workload != "config5":
            extra["generate_audio"] = generate_audio_leg(desc, sd, B, T, K, dev, args.dtype)

    if rank == 0:
        pk = _peaks()
        sb = step_bytes(desc, B, T, cfg=True, elem_bytes=2 if args.dtype == "bf16" else 4)
        alg = sb["total_bytes"]
        ach = alg / (ms_per_step / 1e3) / 1e9
        prof = _profile_facts(args.workload) if (args.batch in (0, WORKLOADS[args.workload]["B"]) and scaling == "weak") else {}
        roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                "traffic": prof.get("dram_bytes_per_step"),
                "kernel": "one sampler step = one CUDA-graph launch of %d kernels; conv_umma_kernel (tcgen05 tap-GEMM) is "
                          "%s of its device time" % (launches_per_step, prof.get("conv_umma_share", "the dominant share")),
                "traffic_source": prof.get("source"),
                "algorithmic_bytes": alg, "weight_bytes": sb["weight_bytes"], "act_bytes_per_row": sb["act_bytes_per_row"],
                "rows": sb["rows"], "peak_source": pk["source"], "flops_per_step": sb["flops"],
                "tensor_frac": sb["flops"] / (ms_per_step / 1e3) / 1e12 / pk["tensor"]}
        cpu, parity, eager = None, None, None
        if world == 1 and not args.no_cpu_baseline:
            parity = parity_check(desc, sd, model, wl, dev)
            cpu = cpu_baseline(desc, sd, wl)
        if world == 1 and not args.no_gpu_eager and not args.quick:
            eager = gpu_eager_baseline(desc, sd, wl, dev)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                "dtype": args.dtype, "data": DATA, "config": _config(wl, world, scaling),
                "e2e": e2e, "gpu_launches": int(launches), "launches_per_step": int(launches_per_step),
                "roofline": roof, "cpu_baseline": cpu, "parity": parity, "gpu_eager_baseline": eager, "extra": extra,
                "clocks": clk}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's per-GPU batch at every N; strong: global batch %d split over N GPUs" % STRONG_GLOBAL_BATCH)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs under ncu)")
    ap.add_argument("--quick", action="store_true", help="skip the ride-along legs (config 2, forward path, GPU eager)")
    ap.add_argument("--batch", type=int, default=0, help="exploratory: override the per-GPU batch of the workload "
                    "(NOT a BASELINE config; the line says so in config.workload)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    wl = dict(WORKLOADS[args.workload])
    note = ""
    if args.scaling == "strong":
        if STRONG_GLOBAL_BATCH % args.gpus != 0:
            raise SystemExit("--scaling strong needs --gpus to divide %d" % STRONG_GLOBAL_BATCH)
        wl["B"] = STRONG_GLOBAL_BATCH // args.gpus
    if args.batch > 0 and args.batch != wl["B"]:
        note = " (exploratory override of %s)" % args.workload
        wl["B"] = args.batch
    wl["name"] = (wl["name"] % (wl["B"], 2 * wl["B"])) + note
    if args.impl == "reference":
        run_reference(args, wl, args.scaling)
    else:
        run_ours(args, wl, args.scaling)


if __name__ == "__main__":
    main()
