"""Host side of the Encodec (SEANet) decoder engine: `EncodecDecoder(...)(latent)` is a drop-in for
`EncodecModel.encodec_model_48khz().decoder(latent)` as the reference calls it (generation.py:34,130).

The arithmetic runs in csrc/codec.cu through the C ABI (`jen1_codec_*`, include/jen1_b200.h); there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from .codec_config import CodecDesc, canonical_state_dict


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


class EncodecCodec:
    """The `codec` object `jen1_b200.generation.Jen1` expects, with the decode side on the B200 engine:
    `decode_latent(latent) -> [B, channels, samples]` (reference generation.py:130).  The encoder side
    (`encode_latent`, reference generation.py:145-150: Encodec encoder + RVQ encode/decode) is not built -- prompts that
    need it pass `init_latent=` instead."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], desc: Optional[CodecDesc] = None, device="cuda:0",
                 precision: str = "tf32"):
        self.decoder = EncodecDecoder(desc, device, precision).load_state_dict(state_dict)
        self.channels = self.decoder.desc.channels
        self.hop = self.decoder.desc.hop

    def decode_latent(self, latent: torch.Tensor) -> torch.Tensor:
        return self.decoder(latent)

    def encode_latent(self, audio: torch.Tensor) -> torch.Tensor:
        raise RuntimeError("jen1_b200: the Encodec ENCODER is not part of this build (pass init_latent= instead)")
