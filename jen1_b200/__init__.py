"""jen1_b200 -- B200-native implementation of the JEN-1 diffusion denoising hot path.

Public surface (mirrors the reference's):
  jen1_b200.generation.Jen1              (reference generation.py:17-132)
  jen1_b200.diffusion.GaussianDiffusion  (reference jen1/diffusion/gdm/gdm.py)
  jen1_b200.model.UNetCFG1d              (reference jen1/model/model.py:268-376)
  jen1_b200.conditioners                 (reference jen1/conditioners.py output contract)
The arithmetic runs in hand-written sm_100a CUDA behind the C ABI in include/jen1_b200.h; importing the model
or engine without the built library (python -m jen1_b200.build) raises -- there is no CPU fallback.
"""
from .config import DiffusionDesc, UNetDesc, latent_frames, tiny_desc  # noqa: F401

__all__ = ["UNetDesc", "DiffusionDesc", "tiny_desc", "latent_frames"]
