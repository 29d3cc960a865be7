"""Encodec (SEANet) decoder engine (csrc/codec.cu) through the C ABI against (a) the committed outputs of the Hugging Face
port of the 48 kHz decoder (tests/golden/codec_decoder.pt; the pip `encodec` package the reference imports --
generation.py:9,34,130 -- is absent offline) and (b) the CPU oracle (oracle/codec_oracle.py) at other shapes.
Tolerances (rel-L2 of the audio): strict mode ("fp32": fp32 FMA everywhere except the recurrent LSTM weights, fp16 in shared
memory with fp32 accumulation) <= 1e-3 (measured 3e-6); default mode ("tf32": conv operands rounded to TF32 for the tensor
cores, fp32 accumulation and storage) <= 1e-2."""
import os

import pytest
import torch

from jen1_b200.codec_config import CodecDesc, random_state_dict, tiny_codec_desc, to_hf_names
from oracle.codec_oracle import decoder_forward

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = {"fp32": 1e-3, "tf32": 1e-2}


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


@pytest.fixture(scope="module", params=["fp32", "tf32"])
def full(request):
    from jen1_b200.codec import EncodecDecoder
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "codec_decoder.pt"))
    desc = CodecDesc()
    sd = random_state_dict(desc, g["weight_seed"])
    return EncodecDecoder(desc, DEV, request.param).load_state_dict(sd), sd, g


def test_decoder_matches_hf_golden(full):
    dec, _, g = full
    assert dec.lstm_cluster() == 16  # hidden 512: 32 units per CTA, W_hh slice 128 KB fp16 in shared memory
    for name, c in g["cases"].items():  # includes T = 3 (shorter than the reflect padding of the k7 conv)
        n0, t0 = dec.launch_count(), dec.tf32_launch_count()
        out = dec(c["z"].to(DEV)).cpu()
        # tensor-core tap-GEMMs: every conv except the narrow last one
        assert dec.tf32_launch_count() - t0 == (0 if dec.precision == "fp32" else 19)
        assert dec.lstm_tc_launch_count() >= 2  # the recurrence ran on the tensor-core cluster kernel
        assert dec.launch_count() - n0 == 24  # pack, 19 tap-GEMMs, the narrow last conv, 2 LSTM cluster launches, final norm
        assert out.shape == c["out"].shape
        assert rel(out, c["out"]) < TOL[dec.precision], (name, rel(out, c["out"]))


# T >= 1024: chunked, overlapped LSTM layers; B = 9: two LSTM clusters (8 sequences + 1)
@pytest.mark.parametrize("B,T", [(1, 1), (2, 77), (3, 150), (1, 700), (2, 1100), (9, 40)])
def test_decoder_matches_oracle_other_shapes(full, B, T):
    dec, sd, _ = full
    z = torch.randn(B, 128, T, generator=torch.Generator().manual_seed(100 + T))
    with torch.no_grad():
        ref = decoder_forward(dec.desc, sd, z)
    n0 = dec.launch_count()
    out = dec(z.to(DEV)).cpu()
    assert dec.launch_count() - n0 == (24 if T < 1024 else 24 + 7 * 3)  # 8 chunks: 7 more launches per LSTM layer + projection
    assert rel(out, ref) < TOL[dec.precision], rel(out, ref)


def test_decoder_batch_rows_are_independent_and_deterministic(full):
    dec, _, _ = full
    z = torch.randn(3, 128, 64, generator=torch.Generator().manual_seed(9)).to(DEV)
    a = dec(z)
    b = dec(z)
    assert torch.equal(a, b)
    alone = dec(z[1:2])
    assert torch.equal(alone[0], a[1])  # a sample's audio does not depend on its batch neighbours (GroupNorm(1) per row)


def test_tiny_decoder_and_hf_key_layout():
    from jen1_b200.codec import EncodecDecoder
    td = tiny_codec_desc()
    sd = random_state_dict(td, 3)
    dec = EncodecDecoder(td, DEV, "fp32").load_state_dict({"decoder." + k: v for k, v in to_hf_names(sd).items()})
    z = torch.randn(2, td.dimension, 37, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = decoder_forward(td, sd, z)
    out = dec(z.to(DEV)).cpu()
    assert out.shape == (2, td.channels, 37 * td.hop)
    assert rel(out, ref) < TOL["fp32"]
    dec2 = EncodecDecoder(td, DEV, "tf32").load_state_dict(sd)
    assert rel(dec2(z.to(DEV)).cpu(), ref) < TOL["tf32"]


def test_decoder_argument_errors(full):
    dec, _, _ = full
    with pytest.raises(ValueError):
        dec(torch.zeros(1, 64, 10))
    with pytest.raises(ValueError):
        dec(torch.zeros(1, 128, 0))
    from jen1_b200.codec import EncodecDecoder
    bad = random_state_dict(CodecDesc(), 0)
    bad.pop("model.15.conv.conv.weight")
    with pytest.raises(KeyError):
        EncodecDecoder(CodecDesc(), DEV).load_state_dict(bad)
