"""Host side of the Encodec (SEANet) decoder engine: `EncodecDecoder(...)(latent)` is a drop-in for
`EncodecModel.encodec_model_48khz().decoder(latent)` as the reference calls it (generation.py:34,130).

The arithmetic runs in csrc/codec.cu through the C ABI (`jen1_codec_*`, include/jen1_b200.h); there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from .codec_config import CodecDesc, canonical_encoder_state_dict, canonical_state_dict


class EncodecDecoder:
    def __init__(self, desc: Optional[CodecDesc] = None, device="cuda:0", precision: str = "tf32"):
        """precision: "tf32" (convs on the TF32 tensor-core tap-GEMM, fp32 storage / accumulation) or "fp32" (strict)."""
        if precision not in ("tf32", "fp32"):
            raise ValueError("precision must be 'tf32' or 'fp32'")
        self.precision = precision
        self.desc = desc or CodecDesc()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("jen1_b200.codec: the decoder engine runs on CUDA only (no CPU fallback)")
        lib = _lib.load()
        d = _lib.Jen1CodecDesc()
        d.channels, d.dimension, d.n_filters = self.desc.channels, self.desc.dimension, self.desc.n_filters
        d.n_ratios = len(self.desc.ratios)
        for i, r in enumerate(self.desc.ratios):
            d.ratios[i] = int(r)
        d.kernel_size, d.last_kernel_size = self.desc.kernel_size, self.desc.last_kernel_size
        d.residual_kernel_size, d.compress, d.lstm_layers = self.desc.residual_kernel_size, self.desc.compress, self.desc.lstm_layers
        d.eps = float(self.desc.eps)
        h = C.c_void_p()
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        rc = lib.jen1_codec_create(C.byref(d), int(index), 1 if precision == "tf32" else 0, C.byref(h))
        if rc != 0:
            raise RuntimeError("jen1_codec_create failed (%d): %s" % (rc, (lib.jen1_codec_last_error(None) or b"").decode()))
        self._lib, self._h, self._finalized = lib, h, False

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.jen1_codec_destroy(h)
            self._h = None

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (what, rc, (self._lib.jen1_codec_last_error(self._h) or b"").decode()))

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        """Decoder tensors in the pip `encodec` layout or the Hugging Face port's (jen1_b200/codec_config.py)."""
        if self._finalized:
            raise RuntimeError("weights are already loaded")
        for name, t in canonical_state_dict(self.desc, sd).items():
            shape = (C.c_int64 * t.dim())(*t.shape)
            self._check(self._lib.jen1_codec_load_tensor(self._h, name.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()),
                        "jen1_codec_load_tensor(%s)" % name)
        self._check(self._lib.jen1_codec_finalize(self._h), "jen1_codec_finalize")
        self._finalized = True
        return self

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    @property
    def hop(self) -> int:
        return self.desc.hop

    def launch_count(self) -> int:
        return int(self._lib.jen1_codec_launch_count(self._h))

    def tf32_launch_count(self) -> int:
        return int(self._lib.jen1_codec_tf32_launch_count(self._h))

    def lstm_tc_launch_count(self) -> int:
        return int(self._lib.jen1_codec_lstm_tc_launch_count(self._h))

    def lstm_cluster(self) -> int:
        return int(self._lib.jen1_codec_lstm_cluster(self._h))

    def workspace_bytes(self, B: int, T: int) -> int:
        return int(self._lib.jen1_codec_workspace_bytes(self._h, int(B), int(T)))

    def __call__(self, latent: torch.Tensor) -> torch.Tensor:
        """latent [B, dimension, T] -> audio [B, channels, T * hop] (fp32, on the engine's device)."""
        if not self._finalized:
            raise RuntimeError("load_state_dict first")
        if latent.dim() != 3 or latent.shape[1] != self.desc.dimension:
            raise ValueError("latent must be [B, %d, T], got %s" % (self.desc.dimension, tuple(latent.shape)))
        if latent.shape[0] < 1 or latent.shape[2] < 1:
            raise ValueError("empty latent")
        z = latent.detach().to(self.device, torch.float32).contiguous()
        B, _, T = z.shape
        out = torch.empty(B, self.desc.channels, T * self.desc.hop, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            self._check(self._lib.jen1_codec_decode(self._h, C.c_void_p(z.data_ptr()), C.c_void_p(out.data_ptr()), int(B), int(T), st),
                        "jen1_codec_decode")
        return out

    forward = __call__


def _fill_desc(desc: CodecDesc):
    d = _lib.Jen1CodecDesc()
    d.channels, d.dimension, d.n_filters = desc.channels, desc.dimension, desc.n_filters
    d.n_ratios = len(desc.ratios)
    for i, r in enumerate(desc.ratios):
        d.ratios[i] = int(r)
    d.kernel_size, d.last_kernel_size = desc.kernel_size, desc.last_kernel_size
    d.residual_kernel_size, d.compress, d.lstm_layers = desc.residual_kernel_size, desc.compress, desc.lstm_layers
    d.eps = float(desc.eps)
    return d


class EncodecEncoder:
    """Encodec (SEANet) ENCODER + residual vector quantizer on the engine: `encode(audio)` is `model.encoder(audio)`
    followed by `quantizer.encode` / `quantizer.decode` (reference generation.py:145-150) for a batch of segments."""

    def __init__(self, desc: Optional[CodecDesc] = None, device="cuda:0", precision: str = "tf32", n_q: int = 16,
                 codebook_size: int = 1024):
        if precision not in ("tf32", "fp32"):
            raise ValueError("precision must be 'tf32' or 'fp32'")
        self.desc, self.precision, self.n_q, self.codebook_size = desc or CodecDesc(), precision, n_q, codebook_size
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("jen1_b200.codec: the encoder engine runs on CUDA only (no CPU fallback)")
        lib = _lib.load()
        d = _fill_desc(self.desc)
        h = C.c_void_p()
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        rc = lib.jen1_codec_create_encoder(C.byref(d), int(index), 1 if precision == "tf32" else 0, int(n_q), int(codebook_size),
                                           C.byref(h))
        if rc != 0:
            raise RuntimeError("jen1_codec_create_encoder failed (%d): %s" % (rc, (lib.jen1_codec_last_error(None) or b"").decode()))
        self._lib, self._h, self._finalized = lib, h, False

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.jen1_codec_destroy(h)
            self._h = None

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (what, rc, (self._lib.jen1_codec_last_error(self._h) or b"").decode()))

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        """The Encodec model's state_dict (pip or Hugging Face key layout): `encoder.*` and the quantizer's codebooks."""
        for name, t in canonical_encoder_state_dict(self.desc, sd, self.n_q, self.codebook_size).items():
            shape = (C.c_int64 * t.dim())(*t.shape)
            self._check(self._lib.jen1_codec_load_tensor(self._h, name.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()),
                        "jen1_codec_load_tensor(%s)" % name)
        self._check(self._lib.jen1_codec_finalize(self._h), "jen1_codec_finalize")
        self._finalized = True
        return self

    def launch_count(self) -> int:
        return int(self._lib.jen1_codec_launch_count(self._h))

    def quantize(self, latent: torch.Tensor):
        """Residual vector quantizer alone: latent [N, dimension, T] -> (codes [n_q, N, T] int32, quantized [N, dimension, T])."""
        z = latent.detach().to(self.device, torch.float32).contiguous()
        N, _, T = z.shape
        codes = torch.empty(self.n_q, N, T, device=self.device, dtype=torch.int32)
        qz = torch.empty_like(z)
        with torch.cuda.device(self.device):
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            self._check(self._lib.jen1_codec_quantize(self._h, C.c_void_p(z.data_ptr()), C.c_void_p(codes.data_ptr()),
                                                      C.c_void_p(qz.data_ptr()), int(N), int(T), st), "jen1_codec_quantize")
        return codes, qz

    def encode(self, audio: torch.Tensor, quantize: bool = True):
        """audio [N, channels, L] -> (latent [N, dimension, T], codes [n_q, N, T] int32, quantized [N, dimension, T]);
        `quantize=False` skips the vector quantizer and returns (latent, None, None)."""
        if not self._finalized:
            raise RuntimeError("load_state_dict first")
        if audio.dim() != 3 or audio.shape[1] != self.desc.channels or audio.shape[0] < 1 or audio.shape[2] < 1:
            raise ValueError("audio must be [N, %d, L], got %s" % (self.desc.channels, tuple(audio.shape)))
        a = audio.detach().to(self.device, torch.float32).contiguous()
        N, _, L = a.shape
        T = -(-L // self.desc.hop)
        lat = torch.empty(N, self.desc.dimension, T, device=self.device, dtype=torch.float32)
        codes = torch.empty(self.n_q, N, T, device=self.device, dtype=torch.int32) if quantize else None
        qz = torch.empty_like(lat) if quantize else None
        with torch.cuda.device(self.device):
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            self._check(self._lib.jen1_codec_encode(self._h, C.c_void_p(a.data_ptr()), C.c_void_p(lat.data_ptr()),
                                                    C.c_void_p(codes.data_ptr()) if quantize else None,
                                                    C.c_void_p(qz.data_ptr()) if quantize else None, int(N), int(L), st),
                        "jen1_codec_encode")
        return lat, codes, qz


class EncodecCodec:
    """The `codec` object `jen1_b200.generation.Jen1` expects, both sides on the B200 engine:
    `decode_latent(latent) -> [B, channels, samples]` (reference generation.py:130) and `encode_latent(audio) ->
    [B, dimension, T]` (reference generation.py:145-150: EncodecModel.encode segment by segment, then quantizer.decode).
    The encode side needs the full Encodec state_dict (`encoder.*`, `quantizer.*`); with a decoder-only state_dict prompts
    that need audio input pass `init_latent=` instead."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], desc: Optional[CodecDesc] = None, device="cuda:0",
                 precision: str = "tf32", sample_rate: int = 48000, segment_s: float = 1.0, overlap: float = 0.01,
                 normalize: bool = True, n_q: int = 16, codebook_size: int = 1024):
        self.decoder = EncodecDecoder(desc, device, precision).load_state_dict(state_dict)
        self.channels = self.decoder.desc.channels
        self.hop = self.decoder.desc.hop
        self.encoder = None
        if any(k.startswith("encoder.") for k in state_dict):  # a full Encodec state_dict: the encode side runs here too
            self.encoder = EncodecEncoder(self.decoder.desc, device, precision, n_q, codebook_size).load_state_dict(state_dict)
        # encodec_model_48khz: 1 s segments, 1 % overlap, per-segment loudness normalisation (encodec/model.py)
        self.segment = int(round(segment_s * sample_rate))
        self.stride = max(1, int((1.0 - overlap) * self.segment))
        self.normalize = normalize

    def decode_latent(self, latent: torch.Tensor) -> torch.Tensor:
        return self.decoder(latent)

    def encode_latent(self, audio: torch.Tensor) -> torch.Tensor:
        """reference generation.py:145-150 (`get_emb`): EncodecModel.encode segment by segment (normalise, encoder, RVQ
        encode), the segments' codes concatenated in time, quantizer.decode -> [B, dimension, T]."""
        if self.encoder is None:
            raise RuntimeError("jen1_b200: this codec was built from a decoder-only state_dict (pass init_latent= instead)")
        B, _, L = audio.shape
        audio = audio.to(self.decoder.device, torch.float32)
        outs = []
        full, offs = [], list(range(0, L, self.stride))
        for off in offs:  # equal-length segments go through the engine as one batch, the ragged tail on its own
            seg = audio[:, :, off: off + self.segment]
            if self.normalize:
                mono = seg.mean(dim=1, keepdim=True)
                scale = mono.pow(2).mean(dim=2, keepdim=True).sqrt() + 1e-8
                seg = seg / scale
            full.append(seg)
        groups, i = [], 0
        while i < len(full):
            j = i
            while j < len(full) and full[j].shape[2] == full[i].shape[2]:
                j += 1
            groups.append((i, j))
            i = j
        for i, j in groups:
            batch = torch.cat(full[i:j], dim=0)  # [(j - i) * B, C, Lseg], segment-major
            lat, _, _ = self.encoder.encode(batch, quantize=False)
            lat = lat.view(j - i, B, lat.shape[1], lat.shape[2])
            outs.extend(lat[s] for s in range(j - i))
        # the quantizer works frame by frame: one pass over the concatenated latents of all segments
        return self.encoder.quantize(torch.cat(outs, dim=2))[1]
