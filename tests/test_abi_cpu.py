"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/jen1_b200.h declares; the product path refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

import jen1_b200
from jen1_b200 import _lib
from jen1_b200.config import UNetDesc, latent_frames, tiny_desc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from jen1_b200 import build
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    header = open(os.path.join(ROOT, "include", "jen1_b200.h")).read()
    declared = set(re.findall(r"\b(jen1_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name


def test_desc_struct_layout_matches_header():
    # 4 + 17 + 16 + 16 + 17 + 8 int32 fields
    assert ctypes.sizeof(_lib.Jen1ModelDesc) == 4 * (4 + 17 + 16 + 16 + 17 + 8)


def test_codec_desc_struct_layout_matches_header():
    # 4 + 8 + 5 int32 fields + one float
    assert ctypes.sizeof(_lib.Jen1CodecDesc) == 4 * (4 + 8 + 5 + 1)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_codec_fails_loudly_without_cuda(lib):
    from jen1_b200.codec import EncodecDecoder
    with pytest.raises(RuntimeError, match="CUDA"):
        EncodecDecoder(device="cpu")
    d = _lib.Jen1CodecDesc()
    h = ctypes.c_void_p()
    assert lib.jen1_codec_create(ctypes.byref(d), 0, 1, ctypes.byref(h)) != 0
    assert b"no CUDA device" in lib.jen1_codec_last_error(None)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_engine_fails_loudly_without_cuda(lib):
    from jen1_b200.engine import Engine, EngineError
    from jen1_b200.weights import random_state_dict
    desc = tiny_desc()
    with pytest.raises(EngineError):
        Engine(desc, random_state_dict(desc, 0), device="cuda:0", dtype="fp32")
    # the raw ABI refuses too
    h = ctypes.c_void_p()
    from jen1_b200.engine import _desc_struct
    ds = _desc_struct(desc)
    assert lib.jen1_engine_create(ctypes.byref(ds), 0, 0, ctypes.byref(h)) != 0
    assert b"no CUDA device" in lib.jen1_last_error(None)


def test_latent_frames():
    assert latent_frames(10) == 1515 and latent_frames(30) == 4545
    assert UNetDesc().level_lengths(4545) == [4545, 4545, 1137, 285, 72, 36, 18, 9, 5, 3]
    assert UNetDesc().level_lengths(1515) == [1515, 1515, 379, 95, 24, 12, 6, 3, 2, 1]


def test_product_does_not_import_oracle():
    import sys
    for mod in ("jen1_b200.engine", "jen1_b200.model", "jen1_b200.diffusion", "jen1_b200.generation"):
        __import__(mod)
    src_dir = os.path.join(ROOT, "jen1_b200")
    for fn in os.listdir(src_dir):
        if fn.endswith(".py"):
            txt = open(os.path.join(src_dir, fn)).read()
            assert "import oracle" not in txt and "from oracle" not in txt, fn
