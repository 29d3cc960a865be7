import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jen1_b200.codec import EncodecDecoder
from jen1_b200.codec_config import CodecDesc, random_state_dict
desc = CodecDesc()
import os as _os
dec = EncodecDecoder(desc, "cuda:0", _os.environ.get("CODEC_PREC", "tf32")).load_state_dict(random_state_dict(desc, 11))
B = int(os.environ.get("CODEC_B", "1"))
for T in [int(a) for a in sys.argv[1:]]:
    z = torch.randn(B, 128, T, device="cuda")
    try:
        out = dec(z); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = int(os.environ.get("CODEC_REPS", "3"))
        each = []
        for _ in range(reps):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(); out = dec(z); a1.record(); torch.cuda.synchronize()
            each.append(a0.elapsed_time(a1))
        if reps > 3:
            print("each:", " ".join("%.1f" % v for v in each), flush=True)
        e0.record()
        for _ in range(3):
            out = dec(z)
        e1.record(); torch.cuda.synchronize()
        print("B", B, "T", T, "ok %.2f ms" % (e0.elapsed_time(e1) / 3), "ws %.2f GB" % (dec.workspace_bytes(B, T) / 1e9), flush=True)
    except Exception as e:
        print("T", T, "FAIL", str(e)[:200], flush=True)
        break
