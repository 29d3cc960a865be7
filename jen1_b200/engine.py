"""Python handle over the C-ABI engine (include/jen1_b200.h).  PyTorch is used only for device memory and
streams: every call hands raw device pointers of torch tensors to the CUDA library on torch's current stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import torch

from . import _lib
from .config import UNetDesc
from .weights import check_state_dict


class EngineError(RuntimeError):
    pass


def _desc_struct(desc: UNetDesc) -> _lib.Jen1ModelDesc:
    s = _lib.Jen1ModelDesc()
    n = desc.num_layers
    if n > _lib.MAX_LEVELS:
        raise ValueError("too many UNet levels")
    s.in_channels, s.out_channels, s.channels, s.num_layers = desc.in_channels, desc.out_channels, desc.channels, n
    for i, m in enumerate(desc.multipliers):
        s.multipliers[i] = m
    for i in range(n):
        s.factors[i] = desc.factors[i]
        s.num_blocks[i] = desc.num_blocks[i]
        s.attentions[i] = desc.attentions[i]
    s.attentions[n] = desc.bottleneck_attention()
    s.resnet_groups = desc.resnet_groups
    s.context_channels = desc.context_channels[0]
    s.context_features_multiplier = desc.context_features_multiplier
    s.context_embedding_features = desc.context_embedding_features
    s.context_embedding_max_length = desc.context_embedding_max_length
    s.attention_heads = desc.attention_heads
    s.attention_multiplier = desc.attention_multiplier
    s.use_skip_scale = int(desc.use_skip_scale)
    return s


class Engine:
    """One engine per GPU.  Not thread-safe.  Fails loudly without CUDA or without the built library."""

    def __init__(self, desc: UNetDesc, state_dict: Dict[str, torch.Tensor], device="cuda:0", dtype: str = "bf16"):
        self._h = None
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise EngineError("jen1_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.desc = desc
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise EngineError("jen1_b200 engine device must be a CUDA device, got %s" % device)
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", self.index)
        self.dtype = dtype
        check_state_dict(desc, state_dict)
        h = C.c_void_p()
        ds = _desc_struct(desc)
        rc = self.lib.jen1_engine_create(C.byref(ds), self.index, {"fp32": 0, "bf16": 1}[dtype], C.byref(h))
        if rc != 0:
            raise EngineError("jen1_engine_create failed (%d): %s" % (rc, self.lib.jen1_last_error(None).decode()))
        self._h = h
        for name, shape, _ in desc.tensor_spec():
            t = state_dict[name].detach().to("cpu", torch.float32).contiguous()
            shp = (C.c_int64 * len(shape))(*shape)
            self._ck(self.lib.jen1_engine_load_tensor(h, name.encode(), C.c_void_p(t.data_ptr()), shp, len(shape)))
        self._ck(self.lib.jen1_engine_finalize(h))
        self.context_epoch = 0  # bumped by every set_context: callers caching "context is current" key on it
        self._t_rows: Dict[int, int] = {}

    # ------------------------------------------------------------------------------------------------
    def _ck(self, rc: int):
        if rc != 0:
            raise EngineError("jen1_b200 engine error: " + self.lib.jen1_last_error(self._h).decode())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _f32(self, t: torch.Tensor, what: str) -> torch.Tensor:
        if t.device != self.device:
            t = t.to(self.device)
        return t.to(torch.float32).contiguous()

    def close(self):
        if getattr(self, "_h", None):
            torch.cuda.synchronize(self.device)
            self.lib.jen1_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------
    def workspace_bytes(self, B: int, T: int) -> int:
        return int(self.lib.jen1_engine_workspace_bytes(self._h, B, T))

    def reserve(self, B: int, T: int):
        self._ck(self.lib.jen1_engine_reserve(self._h, B, T))

    def set_context(self, embedding: torch.Tensor, mask: Optional[torch.Tensor]):
        """embedding [B,S,E]; mask [B,S] (bool or float, True/1 = keep) or None (reference conditioners.py:84-111)."""
        emb = self._f32(embedding, "embedding")
        B, S, E = emb.shape
        if E != self.desc.context_embedding_features:
            raise EngineError("embedding feature size %d != %d" % (E, self.desc.context_embedding_features))
        m = None if mask is None else self._f32(mask, "mask")
        if m is not None and tuple(m.shape) != (B, S):
            raise EngineError("embedding_mask must be [B, S]")
        self._ck(self.lib.jen1_engine_set_context(self._h, C.c_void_p(emb.data_ptr()),
                                                  C.c_void_p(m.data_ptr()) if m is not None else None, B, S,
                                                  self._stream()))
        self._keep = (emb, m)
        self.context_epoch += 1

    def set_timesteps(self, ts: Sequence[int]):
        ts = [int(v) for v in ts]
        arr = (C.c_int64 * len(ts))(*ts)
        self._ck(self.lib.jen1_engine_set_timesteps(self._h, arr, len(ts), self._stream()))
        self._t_rows = {}
        for i, v in enumerate(ts):
            self._t_rows.setdefault(v, i)

    def rows_for(self, ts: Sequence[int]):
        """Conditioning-table rows for the given timesteps, rebuilding the table if some are missing."""
        ts = [int(v) for v in ts]
        if any(v not in self._t_rows for v in ts):
            self.set_timesteps(sorted(set(ts)))
        return [self._t_rows[v] for v in ts]

    def forward(self, x, concat_cond, rows, drop=None, causal=False, embedding_scale=1.0, scale_cfg=False,
                scale_phi=0.7) -> torch.Tensor:
        x = self._f32(x, "x")
        cc = self._f32(concat_cond, "channels")
        B, Cin, T = x.shape
        if Cin != self.desc.in_channels:
            raise EngineError("x has %d channels, expected %d" % (Cin, self.desc.in_channels))
        if tuple(cc.shape) != (B, self.desc.context_channels[0], T):
            raise EngineError("Expected context with %d channels at idx 0" % self.desc.context_channels[0])
        out = torch.empty((B, self.desc.out_channels, T), device=self.device, dtype=torch.float32)
        rows_a = (C.c_int32 * B)(*[int(r) for r in rows])
        dptr = None
        if drop is not None:
            drop = drop.to(self.device).reshape(B).to(torch.uint8).contiguous()
            dptr = C.c_void_p(drop.data_ptr())
        self._ck(self.lib.jen1_unet_forward(self._h, C.c_void_p(x.data_ptr()), C.c_void_p(cc.data_ptr()), rows_a,
                                            dptr, B, T, int(bool(causal)), float(embedding_scale), int(bool(scale_cfg)),
                                            float(scale_phi), C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def sample_begin(self, coef: torch.Tensor, concat_cond: torch.Tensor, B: int, T: int, causal: bool,
                     embedding_scale: float, scale_cfg: bool, scale_phi: float, objective: str, use_graph: bool):
        coef = coef.detach().to("cpu", torch.float32).contiguous()
        S = coef.shape[0]
        assert coef.shape == (S, 8)
        cc = self._f32(concat_cond, "channels")
        if tuple(cc.shape) != (B, self.desc.context_channels[0], T):
            raise EngineError("Expected context with %d channels at idx 0" % self.desc.context_channels[0])
        self._ck(self.lib.jen1_sample_begin(self._h, C.cast(C.c_void_p(coef.data_ptr()), C.POINTER(C.c_float)), S,
                                            C.c_void_p(cc.data_ptr()), B, T, int(bool(causal)), float(embedding_scale),
                                            int(bool(scale_cfg)), float(scale_phi), _lib.OBJECTIVES[objective],
                                            int(bool(use_graph)), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()  # coef is pageable host memory

    def sample_step(self, step: int, x: torch.Tensor, noise: Optional[torch.Tensor], drop: Optional[torch.Tensor]):
        assert x.dtype == torch.float32 and x.is_contiguous() and x.device == self.device
        nptr = None
        if noise is not None:
            assert noise.dtype == torch.float32 and noise.is_contiguous() and noise.shape == x.shape
            nptr = C.c_void_p(noise.data_ptr())
        dptr = None
        if drop is not None:
            assert drop.dtype in (torch.uint8, torch.bool) and drop.is_contiguous() and drop.device == self.device
            dptr = C.c_void_p(drop.data_ptr())
        self._ck(self.lib.jen1_sample_step(self._h, int(step), C.c_void_p(x.data_ptr()), nptr, dptr, self._stream()))

    def attention(self, qkv: torch.Tensor, heads: int, causal: bool = False, impl: str = "tcgen05") -> torch.Tensor:
        """The attention core alone (reference blocks.py:355-380) on a packed bf16 [B, N, 3*C] (q | k | v) tensor."""
        assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and qkv.device == self.device and qkv.dim() == 3
        B, N, C3 = qkv.shape
        ch = C3 // 3
        assert ch * 3 == C3 and ch % heads == 0
        out = torch.empty((B, N, ch), device=self.device, dtype=torch.bfloat16)
        self._ck(self.lib.jen1_attention_forward(self._h, C.c_void_p(qkv.data_ptr()), C.c_void_p(out.data_ptr()), B, N, heads,
                                                 ch // heads, int(bool(causal)), {"tcgen05": 0, "fma": 1, "flash": 2}[impl],
                                                 self._stream()))
        return out

    # ------------------------------------------------------------------------------------------------
    def launch_count(self) -> int:
        return int(self.lib.jen1_engine_launch_count(self._h))

    def umma_launch_count(self) -> int:
        return int(self.lib.jen1_engine_umma_launch_count(self._h))

    def umma_attn_launch_count(self) -> int:
        return int(self.lib.jen1_engine_umma_attn_launch_count(self._h))

    def fused_transformer_launch_count(self) -> int:
        return int(self.lib.jen1_engine_fused_transformer_launch_count(self._h))

    def weight_bytes(self) -> int:
        return int(self.lib.jen1_engine_weight_bytes(self._h))

    def debug_tensor(self, name: str, max_elems: int = 1 << 28) -> torch.Tensor:
        """Intermediate activation of the last `forward` as fp32 [B, L, C] (channels-last), for tests."""
        buf = torch.empty(max_elems if max_elems < (1 << 24) else (1 << 24), dtype=torch.float32)
        shp = (C.c_int64 * 3)()
        rc = self.lib.jen1_engine_debug_tensor(self._h, name.encode(), C.c_void_p(buf.data_ptr()), buf.numel(), shp)
        if rc != 0:
            n = int(shp[0] * shp[1] * shp[2])
            if n > buf.numel():
                buf = torch.empty(n, dtype=torch.float32)
                rc = self.lib.jen1_engine_debug_tensor(self._h, name.encode(), C.c_void_p(buf.data_ptr()), n, shp)
        self._ck(rc)
        n = int(shp[0] * shp[1] * shp[2])
        return buf[:n].reshape(int(shp[0]), int(shp[1]), int(shp[2])).clone()
