#!/usr/bin/env python
"""Encodec decoder engine vs the committed golden (HF port outputs) and the CPU oracle.  Usage: python scripts/codec_check.py [T B]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jen1_b200.codec import EncodecDecoder  # noqa: E402
from jen1_b200.codec_config import CodecDesc, random_state_dict, tiny_codec_desc  # noqa: E402
from oracle.codec_oracle import decoder_forward  # noqa: E402


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def main():
    g = torch.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "codec_decoder.pt"))
    desc = CodecDesc()
    sd = random_state_dict(desc, g["weight_seed"])
    for prec in ("fp32", "tf32"):
        dec = EncodecDecoder(desc, "cuda:0", prec).load_state_dict(sd)
        print(prec, "lstm cluster", dec.lstm_cluster())
        for name, c in g["cases"].items():
            out = dec(c["z"].cuda()).cpu()
            print("golden %-8s rel-L2 %.3e  max abs %.3e" % (name, rel(out, c["out"]), (out - c["out"]).abs().max().item()))
    td = tiny_codec_desc()
    tsd = random_state_dict(td, 3)
    tdec = EncodecDecoder(td, "cuda:0", "tf32").load_state_dict(tsd)
    z = torch.randn(3, td.dimension, 37, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = decoder_forward(td, tsd, z)
    out = tdec(z.cuda()).cpu()
    print("tiny vs oracle rel-L2 %.3e (cluster %d)" % (rel(out, ref), tdec.lstm_cluster()))
    if len(sys.argv) > 2:
        T, B = int(sys.argv[1]), int(sys.argv[2])
        z = torch.randn(B, desc.dimension, T, generator=torch.Generator().manual_seed(2)).cuda()
        out = dec(z)
        torch.cuda.synchronize()
        t = time.time()
        n0 = dec.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = dec(z)
        e1.record()
        torch.cuda.synchronize()
        print("decode B=%d T=%d: %.2f ms (%d launches, workspace %.2f GB), finite %s" %
              (B, T, e0.elapsed_time(e1), dec.launch_count() - n0, dec.workspace_bytes(B, T) / 1e9, bool(torch.isfinite(out).all())))
        if T <= 200:
            with torch.no_grad():
                ref = decoder_forward(desc, sd, z.cpu())
            print("  vs oracle rel-L2 %.3e" % rel(out.cpu(), ref))


if __name__ == "__main__":
    main()
