#!/usr/bin/env python
"""Small workloads for compute-sanitizer / stress runs (scripts/sanitize.sh).

    python scripts/sanitize_driver.py tiny        # tiny UNet: bf16 + fp32 CFG forward (B=2, T=50) and a 3-step DDIM
    python scripts/sanitize_driver.py full        # full-size UNet: one bf16 CFG forward at B=1, T=47 (every tcgen05 plan class)
    python scripts/sanitize_driver.py stress [N]  # determinism stress of the 16-CTA split-K cluster exchange: N forwards
                                                  # (default 500) of the full model at B=2, T=1 -- every conv of the deep
                                                  # levels runs as a (1,1,16) cluster -- each compared BIT-EXACTLY with the first
    python scripts/sanitize_driver.py codec       # Encodec decoder engine: 48 kHz configuration at T=6 + tiny configuration
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jen1_b200.config import UNetDesc, tiny_desc  # noqa: E402
from jen1_b200.diffusion import create_gaussian_diffusion  # noqa: E402
from jen1_b200.model import UNetCFG1d  # noqa: E402
from jen1_b200.weights import random_state_dict  # noqa: E402

DEV = "cuda:0"
KW = dict(embedding_scale=0.8, batch_cfg=True, scale_cfg=True, embedding_mask_proba=0.0)


def inputs(desc, B, T, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, desc.in_channels, T, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    emb = torch.randn(B, desc.context_embedding_max_length, desc.context_embedding_features, generator=g)
    mask = torch.ones(B, desc.context_embedding_max_length, dtype=torch.bool)
    mask[:, -3:] = False
    emb = emb * mask.unsqueeze(-1)
    cc = torch.randn(B, desc.context_channels[0], T, generator=g)
    return [v.to(DEV) for v in (x, t, emb, mask, cc)]


def fwd(model, x, t, emb, mask, cc, **kw):
    y = model(x, t, embedding=emb, embedding_mask=mask, features=None, channels_list=[cc], **dict(KW, **kw))
    torch.cuda.synchronize()
    return y


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    if mode == "tiny":
        desc = tiny_desc()
        sd = random_state_dict(desc, 7)
        for dt in ("bf16", "fp32"):
            m = UNetCFG1d(desc, device=DEV, dtype=dt).load_state_dict(sd)
            a = inputs(desc, 2, 50, 3)
            y = fwd(m, *a)
            yc = fwd(m, *a, causal=True)
            assert torch.isfinite(y).all() and torch.isfinite(yc).all()
            d = create_gaussian_diffusion(steps=1000, noise_schedule="linear", objective="noise", device=DEV,
                                          cfg_dropout_proba=0.2, embedding_scale=0.8, batch_cfg=True, scale_cfg=True,
                                          sampling_steps=3)
            cond = dict(cross_attn_cond=a[2], cross_attn_masks=a[3], global_cond=None, input_concat_cond=a[4])
            out = d.sample(m, (2, desc.in_channels, 50), cond)
            torch.cuda.synchronize()
            assert torch.isfinite(out).all()
            print("tiny %s ok, launches %d" % (dt, m.engine.launch_count()))
    elif mode == "full":
        desc = UNetDesc()
        m = UNetCFG1d(desc, device=DEV, dtype="bf16").load_state_dict(random_state_dict(desc, 0))
        y = fwd(m, *inputs(desc, 1, 47, 5))
        assert torch.isfinite(y).all()
        print("full bf16 T=47 ok, launches %d (tcgen05 conv %d, tcgen05 attention %d)"
              % (m.engine.launch_count(), m.engine.umma_launch_count(), m.engine.umma_attn_launch_count()))
    elif mode == "stress":
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 500
        desc = UNetDesc()
        m = UNetCFG1d(desc, device=DEV, dtype="bf16").load_state_dict(random_state_dict(desc, 0))
        a = inputs(desc, 2, 1, 9)
        ref = fwd(m, *a).clone()
        u0 = m.engine.umma_launch_count()
        t0 = time.time()
        bad = 0
        for i in range(n):
            y = fwd(m, *a)
            if not torch.equal(y, ref):
                bad += 1
        convs = m.engine.umma_launch_count() - u0
        print("split-K determinism stress: %d forwards, %d tcgen05 conv launches (cluster split-K at every deep level), "
              "%d mismatching outputs, %.1f s" % (n, convs, bad, time.time() - t0))
        assert bad == 0
    elif mode == "codec":
        # Encodec decoder engine: the full 48 kHz configuration (16-CTA LSTM cluster, tensor-core kernels) at T = 6 and the
        # tiny configuration in both precisions
        from jen1_b200.codec import EncodecDecoder
        from jen1_b200.codec_config import CodecDesc, random_state_dict as codec_sd, tiny_codec_desc
        for desc, T, B in ((CodecDesc(), 6, 2), (tiny_codec_desc(), 9, 3)):
            sd = codec_sd(desc, 1)
            for prec in ("tf32", "fp32"):
                dec = EncodecDecoder(desc, DEV, prec).load_state_dict(sd)
                out = dec(torch.randn(B, desc.dimension, T, generator=torch.Generator().manual_seed(2)).to(DEV))
                torch.cuda.synchronize()
                assert torch.isfinite(out).all()
                print("codec H=%d %s ok, launches %d (tf32 gemm %d, tensor-core lstm %d)"
                      % (desc.hidden, prec, dec.launch_count(), dec.tf32_launch_count(), dec.lstm_tc_launch_count()))
        from jen1_b200.codec import EncodecEncoder
        from jen1_b200.codec_config import random_encoder_state_dict
        enc = EncodecEncoder(CodecDesc(), DEV, "tf32").load_state_dict(random_encoder_state_dict(CodecDesc(), 2))
        lat, codes, qz = enc.encode(torch.randn(2, 2, 2000, generator=torch.Generator().manual_seed(3)).to(DEV))
        torch.cuda.synchronize()
        assert torch.isfinite(lat).all() and torch.isfinite(qz).all() and int(codes.max()) < 1024
        print("codec encoder + RVQ ok, launches %d" % enc.launch_count())
    else:
        raise SystemExit("unknown mode " + mode)


if __name__ == "__main__":
    main()
