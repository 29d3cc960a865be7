"""Conditioner front-end with the reference's output contract (jen1/conditioners.py:13-208).

Only the *output contract* matters to the denoiser hot path:
    MultiConditioner(batch_metadata, device) -> {id: (emb[B, S, D] fp32, mask[B, S] bool)}
with padded rows of `emb` zeroed (reference conditioners.py:109).  The T5 encoder runs once per generate(),
outside the sampling loop, so it stays host-framework code (HF transformers); its pretrained weights are not
available offline, therefore `EmbeddingConditioner` (caller-supplied embeddings) and `RandomTextConditioner`
(deterministic pseudo-embeddings with the T5 contract, used by benchmarks/tests) are provided beside it.
"""
from __future__ import annotations

import hashlib
import math
from typing import Any, Dict, List, Sequence, Tuple, Union

import torch


class Conditioner:
    """Base: holds the (dim, output_dim, cond_len) triple like reference conditioners.py:13-29."""

    def __init__(self, dim: int, output_dim: int, cond_len: int):
        self.dim, self.output_dim, self.cond_len = dim, output_dim, cond_len

    def __call__(self, inputs: Sequence[Any], device) -> Tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError


class EmbeddingConditioner(Conditioner):
    """Pass-through for precomputed text embeddings: inputs are (emb[S, D], mask[S]) pairs or emb tensors."""

    def __init__(self, output_dim: int = 1024, max_length: int = 128):
        super().__init__(output_dim, output_dim, max_length)

    def __call__(self, inputs, device):
        embs, masks = [], []
        for item in inputs:
            emb, mask = item if isinstance(item, (tuple, list)) else (item, None)
            emb = torch.as_tensor(emb, dtype=torch.float32)
            S = emb.shape[0]
            assert S <= self.cond_len and emb.shape[1] == self.output_dim
            mask = torch.ones(S, dtype=torch.bool) if mask is None else torch.as_tensor(mask).to(torch.bool)
            pad = self.cond_len - S
            if pad:
                emb = torch.cat([emb, emb.new_zeros(pad, emb.shape[1])])
                mask = torch.cat([mask, mask.new_zeros(pad)])
            embs.append(emb * mask.unsqueeze(-1).float())
            masks.append(mask)
        return torch.stack(embs).to(device), torch.stack(masks).to(device)


class RandomTextConditioner(Conditioner):
    """Deterministic stand-in with the T5Conditioner contract: one pseudo-token per whitespace word (max 128),
    N(0,1) features seeded by the text, padded rows zeroed.  NOT a language model -- synthetic conditioning."""

    def __init__(self, output_dim: int = 1024, max_length: int = 128):
        super().__init__(output_dim, output_dim, max_length)

    def __call__(self, texts: List[str], device):
        embs, masks = [], []
        for text in texts:
            n = max(1, min(self.cond_len, len(str(text).split()) + 1))
            seed = int.from_bytes(hashlib.sha256(str(text).encode()).digest()[:8], "little") % (2 ** 63)
            g = torch.Generator().manual_seed(seed)
            emb = torch.zeros(self.cond_len, self.output_dim)
            emb[:n] = torch.randn(n, self.output_dim, generator=g)
            mask = torch.zeros(self.cond_len, dtype=torch.bool)
            mask[:n] = True
            embs.append(emb)
            masks.append(mask)
        return torch.stack(embs).to(device), torch.stack(masks).to(device)


class T5Conditioner(Conditioner):
    """HF T5 encoder conditioner (reference conditioners.py:32-111): tokenise with padding to `max_length`,
    encode, optional projection, zero the padded rows.  Needs locally available pretrained files."""

    DIMS = {"t5-small": 512, "t5-base": 768, "t5-large": 1024, "t5-3b": 1024, "t5-11b": 1024,
            "google/flan-t5-small": 512, "google/flan-t5-base": 768, "google/flan-t5-large": 1024,
            "google/flan-t5-xl": 2048, "google/flan-t5-xxl": 4096}

    def __init__(self, output_dim: int, t5_model_name: str = "t5-base", max_length: int = 128,
                 enable_grad: bool = False, project_out: bool = False):
        assert t5_model_name in self.DIMS, f"Unknown T5 model name: {t5_model_name}"
        dim = self.DIMS[t5_model_name]
        super().__init__(dim, output_dim, max_length)
        from transformers import AutoTokenizer, T5EncoderModel
        self.tokenizer = AutoTokenizer.from_pretrained(t5_model_name)
        self.model = T5EncoderModel.from_pretrained(t5_model_name).train(enable_grad).requires_grad_(enable_grad)
        self.proj_out = torch.nn.Linear(dim, output_dim) if (dim != output_dim or project_out) else torch.nn.Identity()
        self.enable_grad = enable_grad

    def __call__(self, texts: List[str], device):
        self.model.to(device)
        self.proj_out.to(device)
        enc = self.tokenizer(texts, truncation=True, max_length=self.cond_len, padding="max_length", return_tensors="pt")
        ids = enc["input_ids"].to(device)
        mask = enc["attention_mask"].to(device).to(torch.bool)
        self.model.eval()
        with torch.set_grad_enabled(self.enable_grad):
            emb = self.model(input_ids=ids, attention_mask=mask)["last_hidden_state"]
        emb = self.proj_out(emb.float())
        return emb * mask.unsqueeze(-1).float(), mask


class IntConditioner(Conditioner):
    """reference conditioners.py:114-132."""

    def __init__(self, output_dim: int, min_val: int = 0, max_val: int = 512):
        super().__init__(output_dim, output_dim, 1)
        self.min_val, self.max_val = min_val, max_val
        self.int_embedder = torch.nn.Embedding(max_val - min_val + 1, output_dim)

    def __call__(self, ints: List[int], device=None):
        v = torch.tensor(ints).to(device).clamp(self.min_val, self.max_val)
        e = self.int_embedder.to(device)(v).unsqueeze(1)
        return [e, torch.ones(e.shape[0], 1).to(device)]


class NumberEmbedder(torch.nn.Module):
    """reference utils/module.py:58-101: [x, sin(2 pi x w), cos(2 pi x w)] (w in R^{dim/2}, learned) -> Linear(dim + 1, features)."""

    def __init__(self, features: int, dim: int = 256):
        super().__init__()
        assert dim % 2 == 0
        self.features = features
        self.weights = torch.nn.Parameter(torch.randn(dim // 2))
        self.linear = torch.nn.Linear(dim + 1, features)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        shape = x.shape
        x = x.reshape(-1, 1).to(self.weights.dtype)
        freqs = x * self.weights.unsqueeze(0) * 2 * math.pi
        f = torch.cat((x, freqs.sin(), freqs.cos()), dim=-1)
        return self.linear(f).view(*shape, self.features)

    def load_reference_state_dict(self, sd: Dict[str, torch.Tensor], prefix: str = "embedder."):
        """Reference key names: `embedder.embedding.0.weights`, `embedder.embedding.1.{weight,bias}`."""
        self.weights.data.copy_(sd[prefix + "embedding.0.weights"])
        self.linear.weight.data.copy_(sd[prefix + "embedding.1.weight"])
        self.linear.bias.data.copy_(sd[prefix + "embedding.1.bias"])
        return self


class NumberConditioner(Conditioner):
    """reference conditioners.py:135-164: floats clamped to [min_val, max_val], normalised to [0, 1], embedded by
    NumberEmbedder; returns [emb[B, 1, D], ones[B, 1]]."""

    def __init__(self, output_dim: int, min_val: float = 0, max_val: float = 1):
        super().__init__(output_dim, output_dim, 1)
        self.min_val, self.max_val = min_val, max_val
        self.embedder = NumberEmbedder(features=output_dim)

    def __call__(self, floats: List[float], device=None):
        v = torch.tensor([float(x) for x in floats]).to(device).clamp(self.min_val, self.max_val)
        v = (v - self.min_val) / (self.max_val - self.min_val)
        e = self.embedder.to(device)(v).unsqueeze(1)
        return [e, torch.ones(e.shape[0], 1).to(device)]


class MultiConditioner:
    """reference conditioners.py:167-208: apply each conditioner to its key of the per-sample metadata dicts."""

    def __init__(self, conditioners: Dict[str, Conditioner], default_keys: Dict[str, str] = {}):
        self.conditioners, self.default_keys = dict(conditioners), dict(default_keys)

    def __call__(self, batch_metadata: List[Dict[str, Any]], device: Union[torch.device, str]) -> Dict[str, Any]:
        out = {}
        for key, cond in self.conditioners.items():
            inputs = []
            for meta in batch_metadata:
                k = key
                if k not in meta:
                    if k in self.default_keys:
                        k = self.default_keys[k]
                    else:
                        raise ValueError(f"Conditioner key {k} not found in batch metadata")
                v = meta[k]
                if isinstance(v, (list, tuple)) and len(v) == 1:  # collate functions wrap singletons
                    v = v[0]
                inputs.append(v)
            out[key] = cond(inputs, device)
        return out

    forward = __call__


def create_multi_conditioner(config=None, *, text_conditioner: Conditioner = None) -> MultiConditioner:
    """reference utils/script_util.py:151-178 with its INTENDED semantics: one conditioner per entry of
    `conditioning_type` (the reference returns from inside the loop, so only 't5' is ever built -- SURVEY.md 3.6).
    `config` is a ConditionerConfig-like namespace (utils/conditioner_config.py:10-37): cond_dim, default_keys,
    conditioning_type and the per-type sub-configs with an `id` (= the metadata key).  `text_conditioner` replaces the
    T5 encoder (its pretrained weights are not reachable offline): e.g. EmbeddingConditioner / RandomTextConditioner."""
    from types import SimpleNamespace
    if config is None:
        config = SimpleNamespace(cond_dim=1024, default_keys={}, conditioning_type=["t5", "int", "number"],
                                 t5_config=SimpleNamespace(id="prompt", t5_model_name="google/flan-t5-large", max_length=128, project_out=True),
                                 int_config=SimpleNamespace(id="seconds_start", min_val=0, max_val=512),
                                 number_config=SimpleNamespace(id="seconds_total", min_val=0, max_val=512))

    def as_dict(c):
        d = {k: getattr(c, k) for k in dir(c) if not k.startswith("_") and not callable(getattr(c, k))}
        return d.pop("id", None), d

    conds: Dict[str, Conditioner] = {}
    for kind in config.conditioning_type:
        if kind == "t5":
            cid, kw = as_dict(config.t5_config)
            conds[cid] = text_conditioner if text_conditioner is not None else T5Conditioner(output_dim=config.cond_dim, **kw)
        elif kind == "int":
            cid, kw = as_dict(config.int_config)
            conds[cid] = IntConditioner(output_dim=config.cond_dim, **kw)
        elif kind == "number":
            cid, kw = as_dict(config.number_config)
            conds[cid] = NumberConditioner(output_dim=config.cond_dim, **kw)
        else:
            raise NotImplementedError("Invalid conditioner type: %s" % kind)
    return MultiConditioner(conds, default_keys=dict(config.default_keys))
