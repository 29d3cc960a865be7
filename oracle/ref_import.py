"""Import shim for the UNMODIFIED reference: /root/reference (build container) or the git-ignored install
baseline/_ref/ made by scripts/install_reference.py (travels to the GPU box).

TEST INFRASTRUCTURE -- never imported by the product path.  Used by `oracle/make_golden.py` (fixture generation
in the build container) and by `bench.py --impl reference` (the reference's own CPU path as the timed baseline).
Follows SURVEY.md Appendix A: two missing third-party modules are stubbed
(`einops_exts.rearrange_many` -- reference jen1/model/blocks.py:8, utils/module.py:7;
`dac.nn.layers.Snake1d` -- blocks.py:5, only constructed when use_snake=True).
"""
import os
import sys
import types

import einops
import torch

REFERENCE_ROOT = "/root/reference"
INSTALLED_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def reference_root(prefer_installed: bool = False):
    """Where the reference can be imported from, or None."""
    order = (INSTALLED_ROOT, REFERENCE_ROOT) if prefer_installed else (REFERENCE_ROOT, INSTALLED_ROOT)
    for r in order:
        if os.path.isdir(os.path.join(r, "jen1", "model")):
            return r
    return None


def install_shims(root=None):
    if "einops_exts" not in sys.modules:
        ee = types.ModuleType("einops_exts")
        ee.rearrange_many = lambda ts, pattern, **kw: tuple(einops.rearrange(t, pattern, **kw) for t in ts)
        sys.modules["einops_exts"] = ee
    if "dac" not in sys.modules:
        dac, dnn, dl = (types.ModuleType(n) for n in ("dac", "dac.nn", "dac.nn.layers"))

        class Snake1d(torch.nn.Module):
            def __init__(self, channels):
                super().__init__()

        dl.Snake1d = Snake1d
        sys.modules.update({"dac": dac, "dac.nn": dnn, "dac.nn.layers": dl})
    root = root or reference_root()
    if root is None:
        raise ImportError("the reference is neither at %s nor installed at %s" % (REFERENCE_ROOT, INSTALLED_ROOT))
    if root not in sys.path:
        sys.path.insert(0, root)


def reference_model_kwargs():
    install_shims()
    from utils.config import Config
    cfg = {k: v for k, v in Config.model_config.__dict__.items() if not k.startswith("__") and not callable(v)}
    return cfg


def build_reference_unet(**overrides):
    """UNetCFG1d(**ModelConfig) exactly as reference utils/script_util.py:271-284 does."""
    install_shims()
    from jen1.model.model import UNetCFG1d
    cfg = reference_model_kwargs()
    cfg.update(overrides)
    cef = cfg.pop("context_embedding_features")
    cml = cfg.pop("context_embedding_max_length")
    return UNetCFG1d(context_embedding_features=cef, context_embedding_max_length=cml, **cfg).eval()


def build_reference_diffusion(sampling_steps=100, **kw):
    install_shims()
    from utils.script_util import create_gaussian_diffusion
    args = dict(steps=1000, noise_schedule="linear", objective="noise", device="cpu", cfg_dropout_proba=0.2,
                embedding_scale=0.8, batch_cfg=True, scale_cfg=True, sampling_steps=sampling_steps)
    args.update(kw)
    return create_gaussian_diffusion(**args)
