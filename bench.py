#!/usr/bin/env python
"""Benchmark of the JEN-1 denoiser hot path (BASELINE.json metric: denoiser latent-frames/sec/step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config2|config3|config5]

One "step" = one DDIM sampler step of the hot path over one batch: pack x -> UNetCFG1d on the doubled CFG batch
-> CFG combine + std rescale -> x0/eps conversion + clamp -> DDIM update (reference gdm.py:202-222), plus the
per-step RNG draws the reference makes (bernoulli cond-dropout, randn_like).  value = B*T latent frames per GPU
* N GPUs / seconds-per-step.

  * `value`  : K steps timed with CUDA events on the launching stream, everything resident in HBM.
  * `e2e`    : the same metric through the public API -- `GaussianDiffusion.sample(model, shape, conditioning)`
               with the conditioning in pinned HOST memory (H2D inside the timed region) and the finished latent
               read back to the host (D2H inside), K sampler steps per call.
  * `roofline`: whole-step HBM roofline (one CUDA-graph launch = one step): algorithmic bytes (jen1_b200.workload)
               / event-timed step duration vs MEASURED_PEAKS.json.
  * `cpu_baseline`: the oracle port of the reference (oracle/, PyTorch CPU fp32) on this box's host cores, same
               workload, a bounded number of sampler steps.
  * `--impl reference`: the reference CPU arm alone (oracle port; the reference itself is pure Python that
               cannot travel to the GPU box -- see DESIGN.md).

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
    "config2": dict(B=1, seconds=10, T=1515, name="100-step DDIM, 10 s (T=1515) latent, batch 1 per GPU, CFG 0.8 (2 UNet rows)"),
    # BASELINE.json configs[2]: 30 s, batch 32 over 8 GPUs = 4 samples per GPU
    "config3": dict(B=4, seconds=30, T=4545, name="100-step DDIM, 30 s (T=4545) latent, batch 4 per GPU, CFG 0.8 (8 UNet rows)"),
    # BASELINE.json configs[4]: continuation (masked latent, causal convs + causal self-attention), 30 s, batch 16 over 4 GPUs
    "config5": dict(B=4, seconds=30, T=4545, causal=True,
                    name="100-step DDIM continuation (causal), 30 s (T=4545) masked latent, batch 4 per GPU, CFG 0.8 (8 UNet rows)"),
}
METRIC = "denoiser latent-frames/sec/step"
UNIT = "latent-frames/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tensor=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor=1590.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return self
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


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


# ------------------------------------------------------------------------------------------------------------
def run_reference(args, wl):
    """Reference CPU arm: the oracle port of UNetCFG1d + DDIM on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from jen1_b200.config import UNetDesc
    from jen1_b200.weights import random_state_dict
    from oracle.gdm_oracle import OracleDiffusion
    from oracle.unet_oracle import OracleUNet
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    desc = UNetDesc()
    B, T = wl["B"], wl["T"]
    sd = random_state_dict(desc, 0)
    model = OracleUNet(desc, sd)
    dif = OracleDiffusion(sampling_timesteps=100)
    causal = bool(wl.get("causal", False))
    emb, mask, cc = _make_problem(desc, B, T, 1, causal)
    cond = dict(cross_attn_cond=emb, cross_attn_masks=mask, global_cond=None, input_concat_cond=cc)
    pairs = dif.time_pairs()
    torch.manual_seed(0)
    x = torch.randn(B, desc.in_channels, T)

    def step(i, x):
        time_, time_next = pairs[i % (len(pairs) - 1)]
        tc = torch.full((B,), time_, dtype=torch.long)
        eps, x0 = dif.model_predictions(x, tc, model, cond, clip=True, causal=causal)
        a, an = dif.alphas_cumprod[time_], dif.alphas_cumprod[time_next]
        sigma = dif.eta * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
        c = (1 - an - sigma ** 2).sqrt()
        return x0 * an.sqrt() + c * eps + sigma * torch.randn_like(x)

    with torch.no_grad():
        for i in range(args.warmup):
            x = step(i, x)
        t0 = time.perf_counter()
        for i in range(args.steps):
            x = step(args.warmup + i, x)
        dt = time.perf_counter() - t0
    ms = dt / args.steps * 1e3
    val = B * T / (ms / 1e3)
    sample = "%d sampler steps of the full workload (B=%d, T=%d, CFG -> %d UNet rows), fp32" % (args.steps, B, T, 2 * B)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (seeded random-init weights, N(0,1) latents/embeddings)",
            "config": {"workload": wl["name"], "frames_per_step": B * T, "device": "host CPU"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline(desc, sd, wl, budget_s=20.0):
    import torch
    from oracle.gdm_oracle import OracleDiffusion
    from oracle.unet_oracle import OracleUNet
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, T = wl["B"], wl["T"]
    model = OracleUNet(desc, sd)
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
            time_, _ = pairs[i % len(pairs)]
            tc = torch.full((B,), time_, dtype=torch.long)
            t0 = time.perf_counter()
            eps, x0 = dif.model_predictions(x, tc, model, cond, clip=True, causal=causal)
            x = x0 * 0.99 + 0.1 * eps + 0.05 * torch.randn_like(x)
            times.append(time.perf_counter() - t0)
            i += 1
            if i >= 3 and (time.perf_counter() - t_start > budget_s or i >= 40):
                break
    warm = sorted(times[2:]) if len(times) > 3 else sorted(times)
    med = warm[len(warm) // 2]
    return {"value": B * T / med, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d sampler steps (median of the last %d) of the full workload B=%d T=%d on the oracle port, fp32"
                      % (len(times), len(warm), B, T), "ms_per_step": med * 1e3}


# ------------------------------------------------------------------------------------------------------------
def run_ours(args, wl):
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
    sd = random_state_dict(desc, 0)
    model = UNetCFG1d(desc, device=dev, dtype=args.dtype).load_state_dict(sd)
    eng = model.engine
    S = max(K + W + 1, 20)  # schedule length (the is-last row is never used): one conditioning row / coefficient row per step
    dif = create_gaussian_diffusion(steps=1000, noise_schedule="linear", objective="noise", device=dev,
                                    cfg_dropout_proba=0.2, embedding_scale=0.8, batch_cfg=True, scale_cfg=True,
                                    sampling_steps=S)
    # each rank owns its own shard of the global batch (distinct prompts): seed by rank
    emb_h, mask_h, cc_h = _make_problem(desc, B, T, 1 + rank, bool(wl.get("causal", False)))
    emb_p, mask_p, cc_p = emb_h.pin_memory(), mask_h.pin_memory(), cc_h.pin_memory()

    # ---------------- device-resident timing: K sampler steps, CUDA events on the launching stream
    side = torch.cuda.Stream(dev)
    with torch.cuda.stream(side):
        emb, mask, cc = emb_p.to(dev), mask_p.to(dev), cc_p.to(dev)
        model.set_context(emb, mask)
        eng.set_timesteps([t for t, _ in dif.time_pairs()])
        coef = dif.ddim_coefficients()
        coef[:, 7] = 0.0  # no "last step" shortcut inside the timed window: every step does the full update
        coef[:, 4:7] = torch.nan_to_num(coef[:, 4:7])
        eng.sample_begin(coef, cc, B, T, bool(wl.get("causal", False)), 0.8, True, 0.7, "noise", True)
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
        launches_per_step = (eng.launch_count() - l0) // max(W, 1)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        clocks = ClockSampler(local).start() if rank == 0 else None
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
        if world > 1:
            dist.barrier()
    clk = clocks.stop() if clocks else None
    t_ms = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = t_ms.item() / K
    value = world * B * T / (ms_per_step / 1e3)

    # ---------------- end to end through the public API with host buffers
    dif_e = create_gaussian_diffusion(steps=1000, noise_schedule="linear", objective="noise", device=dev,
                                      cfg_dropout_proba=0.2, embedding_scale=0.8, batch_cfg=True, scale_cfg=True,
                                      sampling_steps=max(K, 20))
    Se = dif_e.sampling_timesteps
    out_h = torch.empty(B, desc.in_channels, T).pin_memory()

    def e2e_call():
        cond = dict(cross_attn_cond=emb_p.to(dev, non_blocking=True), cross_attn_masks=mask_p.to(dev, non_blocking=True),
                    global_cond=None, input_concat_cond=cc_p.to(dev, non_blocking=True))
        lat = dif_e.sample(model, (B, desc.in_channels, T), cond, causal=bool(wl.get("causal", False)))
        out_h.copy_(lat, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    e2e = None
    if not args.no_e2e:
        e2e_call()  # warm (graph capture for this shape, allocator)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        e2e_call()
        torch.cuda.synchronize(dev)
        e_ms = (time.perf_counter() - t0) * 1e3
        t_e = torch.tensor([e_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        e_ms_step = t_e.item() / Se
        h2d = (emb_p.numel() * 4 + mask_p.numel() + cc_p.numel() * 4)
        d2h = out_h.numel() * 4
        e2e = {"value": world * B * T / (e_ms_step / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d / Se,
               "d2h_bytes_per_step": d2h / Se, "steps_per_call": Se, "ms_per_step": e_ms_step,
               "api": "GaussianDiffusion.sample(UNetCFG1d, shape, conditioning): pinned-host conditioning in, host latent out; "
                      "includes the per-call context K/V hoist and timestep tables"}

    if rank == 0:
        pk = _peaks()
        sb = step_bytes(desc, B, T, cfg=True, elem_bytes=2 if args.dtype == "bf16" else 4)
        alg = sb["total_bytes"]
        ach = alg / (ms_per_step / 1e3) / 1e9
        prof = _profile_facts(args.workload) if args.batch in (0, WORKLOADS[args.workload]["B"]) else {}
        roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                "traffic": prof.get("dram_bytes_per_step"),
                "kernel": "one sampler step = one CUDA-graph launch of %d kernels; conv_umma_kernel (tcgen05 tap-GEMM) is "
                          "%s of its device time" % (launches_per_step, prof.get("conv_umma_share", "the dominant share")),
                "traffic_source": prof.get("source"),
                "algorithmic_bytes": alg, "weight_bytes": sb["weight_bytes"], "act_bytes_per_row": sb["act_bytes_per_row"],
                "rows": sb["rows"], "peak_source": pk["source"], "flops_per_step": sb["flops"],
                "tensor_frac": sb["flops"] / (ms_per_step / 1e3) / 1e12 / pk["tensor"]}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(desc, sd, wl)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.dtype, "data": "synthetic (seeded random-init weights, N(0,1) latents/embeddings)",
                "config": {"workload": wl["name"], "frames_per_step_per_gpu": B * T, "parallelism": "dp%d (batch sharding, no collective)" % world,
                           "l2": "per-step working set (%.0f MB weights + activations) exceeds the 126 MB L2" % (alg / 1e6)},
                "e2e": e2e, "gpu_launches": int(launches), "launches_per_step": int(launches_per_step),
                "roofline": roof, "cpu_baseline": cpu, "clocks": clk}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs under ncu)")
    ap.add_argument("--batch", type=int, default=0, help="exploratory: override the per-GPU batch of the workload "
                    "(NOT a BASELINE config; the line says so in config.workload)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    wl = dict(WORKLOADS[args.workload])
    if args.batch > 0 and args.batch != wl["B"]:
        wl["name"] = wl["name"].replace("batch %d per GPU" % wl["B"], "batch %d per GPU (exploratory override of %s)" % (args.batch, args.workload))
        wl["name"] = wl["name"].replace("(%d UNet rows)" % (2 * wl["B"]), "(%d UNet rows)" % (2 * args.batch))
        wl["B"] = args.batch
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
