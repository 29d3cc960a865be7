import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jen1_b200.codec import EncodecCodec
from jen1_b200.codec_config import CodecDesc, random_encoder_state_dict, random_state_dict
desc = CodecDesc()
sd = dict(random_state_dict(desc, 11)); sd.update(random_encoder_state_dict(desc, 21))
codec = EncodecCodec(sd, desc, "cuda:0")
audio = (torch.randn(4, 2, 1440000, generator=torch.Generator().manual_seed(5)) * 0.3).cuda()
for _ in range(2):
    lat = codec.encode_latent(audio)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); lat = codec.encode_latent(audio); e1.record(); torch.cuda.synchronize()
print("encode 4 x 30 s: %.2f ms, latent %s" % (e0.elapsed_time(e1), tuple(lat.shape)))
