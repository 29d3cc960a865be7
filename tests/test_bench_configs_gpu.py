"""GPU parity of the bf16 tcgen05 engine -- the one bench.py measures -- at the BENCHMARKED configurations, against the
CPU oracle (oracle/, pinned to the live reference) and the committed reference goldens.  Through the C ABI.

Tolerances (SURVEY.md section 7 "hard parts" 3, calibrated on the reference's own bf16-autocast run: 8.9e-3 for one
evaluation, 2.8e-2 after a 20-step DDIM):
  one UNet evaluation, bf16 storage / fp32 accumulate : rel-L2 <= 1e-2
  100-step DDIM trajectory (eta = 1, clamp)            : rel-L2 <= 5e-2 on the final latent
"""
import os

import pytest
import torch

from jen1_b200.config import UNetDesc, tiny_desc
from jen1_b200.weights import random_state_dict
from oracle.make_golden import VARIANTS, make_inputs
from oracle.unet_oracle import OracleUNet, unet_cfg_forward

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL_EVAL = 1e-2
TOL_TRAJ = 5e-2
CFG = dict(embedding_scale=0.8, batch_cfg=True, scale_cfg=True, embedding_mask_proba=0.0)


def rel_l2(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def full_bf16():
    from jen1_b200.model import UNetCFG1d
    desc = UNetDesc()
    sd = random_state_dict(desc, 0)
    model = UNetCFG1d(desc, device=DEV, dtype="bf16").load_state_dict(sd)
    torch.set_num_threads(os.cpu_count() or 1)
    return desc, sd, model


def _engine(model, x, t, emb, mask, cc, **kw):
    y = model(x.to(DEV), t.to(DEV), embedding=emb.to(DEV), embedding_mask=mask.to(DEV), features=None,
              channels_list=[cc.to(DEV)], **kw)
    torch.cuda.synchronize()
    return y.cpu()


def _oracle(desc, sd, x, t, emb, mask, cc, **kw):
    with torch.no_grad():
        return unet_cfg_forward(desc, sd, x, t, embedding=emb, embedding_mask=mask, channels_list=[cc], **kw)


def test_config2_shape_matches_reference_golden_and_oracle(full_bf16, golden_dir):
    """BASELINE configs[1]: B=1, T=1515, CFG (2 UNet rows).  Golden = output of the UNMODIFIED reference."""
    desc, sd, model = full_bf16
    fx = torch.load(os.path.join(golden_dir, "unet_full_c2.pt"))
    rec = fx["cases"]["T1515_B1"]
    x, t, emb, mask, cc = make_inputs(desc, rec["B"], rec["T"], rec["seed"], rec["masked_tail"])
    y = _engine(model, x, t, emb, mask, cc, **CFG)
    assert model.engine.umma_launch_count() > 0
    assert model.engine.umma_attn_launch_count() + model.engine.fused_transformer_launch_count() > 0
    e_ref = rel_l2(y, rec["outputs"]["cfg"])
    print("config2 bf16 vs reference golden: rel-L2 %.3e" % e_ref)
    assert e_ref < TOL_EVAL, e_ref
    # the bench's own conditioning: zero masked latent + zero mask (text-guided)
    cc0 = torch.zeros_like(cc)
    e_or = rel_l2(_engine(model, x, t, emb, mask, cc0, **CFG), _oracle(desc, sd, x, t, emb, mask, cc0, **CFG))
    print("config2 bf16 vs oracle (zero concat cond): rel-L2 %.3e" % e_or)
    assert e_or < TOL_EVAL, e_or


def _config3_inputs(desc, seed, continuation):
    """4 samples at T=4545 with four distinct embeddings / timesteps, one with a masked (zeroed) context tail; concat
    conditioning as bench.py builds it (zeros for text-guided; masked stand-in latent + keep mask for continuation)."""
    B, T = 4, 4545
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, desc.in_channels, T, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    emb = torch.randn(B, desc.context_embedding_max_length, desc.context_embedding_features, generator=g)
    mask = torch.ones(B, desc.context_embedding_max_length, dtype=torch.bool)
    mask[2, -60:] = False
    emb = emb * mask.unsqueeze(-1)
    cc = torch.zeros(B, desc.context_channels[0], T)
    if continuation:
        lat = torch.randn(B, desc.in_channels, T, generator=g) * 0.5
        keep = torch.zeros(B, 1, T)
        keep[:, :, : T // 2] = 1.0
        cc = torch.cat([lat * keep, keep], dim=1)
    return x, t, emb, mask, cc


def test_config3_shape_matches_oracle(full_bf16):
    """BASELINE configs[2] per-GPU shard: 4 samples x T=4545, CFG -> 8 UNet rows (tiles fold batch rows, multi-slot
    GroupNorm statistics, split-K clusters spanning rows)."""
    desc, sd, model = full_bf16
    x, t, emb, mask, cc = _config3_inputs(desc, 31, False)
    y = _engine(model, x, t, emb, mask, cc, **CFG)
    ref = _oracle(desc, sd, x, t, emb, mask, cc, **CFG)
    errs = [rel_l2(y[i], ref[i]) for i in range(4)]
    print("config3 bf16 vs oracle: rel-L2 %.3e, per sample %s" % (rel_l2(y, ref), ["%.2e" % e for e in errs]))
    assert rel_l2(y, ref) < TOL_EVAL and max(errs) < TOL_EVAL, errs


def test_config5_shape_causal_masked_concat_matches_oracle(full_bf16):
    """BASELINE configs[4] per-GPU shard (continuation): causal convs + causal self-attention, masked stand-in latent and
    keep mask as input-concat conditioning, 4 samples x T=4545."""
    desc, sd, model = full_bf16
    x, t, emb, mask, cc = _config3_inputs(desc, 32, True)
    kw = dict(CFG, causal=True)
    y = _engine(model, x, t, emb, mask, cc, **kw)
    ref = _oracle(desc, sd, x, t, emb, mask, cc, **kw)
    errs = [rel_l2(y[i], ref[i]) for i in range(4)]
    print("config5 bf16 vs oracle: rel-L2 %.3e, per sample %s" % (rel_l2(y, ref), ["%.2e" % e for e in errs]))
    assert rel_l2(y, ref) < TOL_EVAL and max(errs) < TOL_EVAL, errs


def test_bf16_100_step_trajectory_matches_oracle(full_bf16):
    """100-step DDIM (eta=1, CFG 0.8, cond-dropout 0.2, clamp) of the full model at config 1's shape: bf16 engine with
    the CUDA-graph step vs the fp32 oracle, identical random stream (CPU generator, reference draw order)."""
    from jen1_b200.diffusion import create_gaussian_diffusion
    from oracle.gdm_oracle import OracleDiffusion
    desc, sd, model = full_bf16
    B, T, S = 1, 150, 100
    x, t, emb, mask, cc = make_inputs(desc, B, T, 301, 0)
    cc = torch.zeros_like(cc)
    cond = dict(cross_attn_cond=emb, cross_attn_masks=mask, global_cond=None, input_concat_cond=cc)
    oracle_model = OracleUNet(desc, sd)
    states = []

    def recording(xs, ts, **kw):  # the state the sampler hands to the model at every step
        states.append(xs.clone())
        return oracle_model(xs, ts, **kw)

    torch.manual_seed(2024)
    ref_final = OracleDiffusion(sampling_timesteps=S).sample(recording, (B, desc.in_channels, T), cond)
    cond_d = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in cond.items()}
    d = create_gaussian_diffusion(steps=1000, noise_schedule="linear", objective="noise", device=DEV,
                                  cfg_dropout_proba=0.2, embedding_scale=0.8, batch_cfg=True, scale_cfg=True,
                                  sampling_steps=S, rng_device="cpu", use_cuda_graph=True)
    torch.manual_seed(2024)
    got = d.sample(model, (B, desc.in_channels, T), cond_d, return_all_timesteps=True).cpu()
    # reference gdm.py:199-225: the stack holds the initial noise twice, then the state before every later step
    assert got.shape[1] == S + 1 and len(states) == S
    per_step = [0.0] + [rel_l2(got[:, i + 1], states[i]) for i in range(S)]
    torch.manual_seed(2024)
    final = d.sample(model, (B, desc.in_channels, T), cond_d).cpu()
    e_final = rel_l2(final, ref_final)
    print("bf16 100-step DDIM vs oracle: final latent rel-L2 %.3e; state after 25/50/75/100 steps %.2e %.2e %.2e %.2e"
          % (e_final, per_step[25], per_step[50], per_step[75], per_step[S]))
    assert e_final < TOL_TRAJ, e_final
    assert max(per_step) < TOL_TRAJ, max(per_step)


def test_cfg_dropout_goldens_with_explicit_mask(full_bf16, golden_dir):
    """The reference's stochastic cond-dropout outputs (model.py:323-328): the bernoulli draw is reproduced on the CPU
    generator exactly as utils/module.py:36-42 makes it and handed to the engine as an explicit drop mask."""
    from jen1_b200.model import UNetCFG1d
    desc, sd, model = full_bf16
    tdesc = tiny_desc()
    tsd = random_state_dict(tdesc, 7)
    tiny32 = UNetCFG1d(tdesc, device=DEV, dtype="fp32").load_state_dict(tsd)
    checked = 0
    for fn, dsc, mdl, tol in (("unet_tiny.pt", tdesc, tiny32, 5e-4), ("unet_full.pt", desc, model, TOL_EVAL),
                              ("unet_full_c2.pt", desc, model, TOL_EVAL)):
        fx = torch.load(os.path.join(golden_dir, fn))
        for name, rec in fx["cases"].items():
            if mdl is model and rec["T"] < 16:
                continue
            x, t, emb, mask, cc = make_inputs(dsc, rec["B"], rec["T"], rec["seed"], rec["masked_tail"])
            for v, ref in rec["outputs"].items():
                if not v.startswith("cfg_dropout"):
                    continue
                seed = int(v.split("seed")[1])
                torch.manual_seed(seed)
                drop = torch.bernoulli(torch.full((rec["B"], 1, 1), 0.5)).to(torch.bool)
                y = _engine(mdl, x, t, emb, mask, cc, embedding_scale=0.8, batch_cfg=True, scale_cfg=True,
                            embedding_mask_proba=0.5, drop_mask=drop.to(DEV))
                err = rel_l2(y, ref)
                assert err < tol, "%s %s %s: rel-L2 %.3e (drop=%s)" % (fn, name, v, err, drop.reshape(-1).tolist())
                checked += 1
    assert checked >= 6


def test_training_losses_through_engine(golden_dir):
    """`training_loosses` (reference gdm.py:245-272) with the ENGINE as the model: q_sample -> UNetCFG1d forward with CFG +
    cond-dropout -> MSE, against the reference's golden loss (same CPU random stream: rand_like noise, then bernoulli)."""
    from jen1_b200.diffusion import create_gaussian_diffusion
    from jen1_b200.model import UNetCFG1d
    fx = torch.load(os.path.join(golden_dir, "gdm.pt"))
    desc = tiny_desc()
    sd = random_state_dict(desc, 7)
    x, t, emb, mask, cc = make_inputs(desc, 3, 50, 31, 0)
    cond = dict(cross_attn_cond=emb.to(DEV), cross_attn_masks=mask.to(DEV), global_cond=None, input_concat_cond=cc.to(DEV))
    for dtype, tol in (("fp32", 2e-4), ("bf16", 2e-2)):
        model = UNetCFG1d(desc, device=DEV, dtype=dtype).load_state_dict(sd)
        d = create_gaussian_diffusion(steps=1000, noise_schedule="linear", objective="noise", device=DEV,
                                      cfg_dropout_proba=0.2, embedding_scale=0.8, batch_cfg=True, scale_cfg=True,
                                      sampling_steps=100, rng_device="cpu")
        torch.manual_seed(fx["train_loss"]["rng_seed"])
        noise = torch.rand_like(x)  # the reference's default noise is UNIFORM (gdm.py:247), drawn before the model call
        loss = float(d.training_loosses(model, x.to(DEV), t.to(DEV), cond, noise=noise.to(DEV), causal=False))
        ref = fx["train_loss"]["loss"]
        assert abs(loss - ref) < tol * max(1.0, abs(ref)), (dtype, loss, ref)


def test_two_engines_on_two_devices_in_one_process():
    """One handle per GPU in the same process (ABI threading contract): per-device kernel attributes must be set for
    every engine.  Needs >= 2 GPUs; skipped otherwise."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from jen1_b200.model import UNetCFG1d
    desc = tiny_desc()
    sd = random_state_dict(desc, 7)
    x, t, emb, mask, cc = make_inputs(desc, 2, 50, 11, 3)
    ref = _oracle(desc, sd, x, t, emb, mask, cc, **CFG)
    for dev in ("cuda:0", "cuda:1"):
        m = UNetCFG1d(desc, device=dev, dtype="bf16").load_state_dict(sd)
        y = m(x.to(dev), t.to(dev), embedding=emb.to(dev), embedding_mask=mask.to(dev), features=None,
              channels_list=[cc.to(dev)], **CFG).cpu()
        assert rel_l2(y, ref) < 2e-2, dev
