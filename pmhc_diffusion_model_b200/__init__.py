"""B200-native denoising hot path of cmbi/pmhc-diffusion-model.

Drop-in surface (same names and argument meaning as the reference's `diffusion` package):

    from pmhc_diffusion_model_b200.diffusion.model import Model
    from pmhc_diffusion_model_b200.diffusion.optimizer import DiffusionModelOptimizer

Everything numerical runs in hand-written sm_100a CUDA kernels behind the C ABI declared in
`include/pmhc_b200.h` (`libpmhc_b200.so`, built by `__graft_entry__.build()`); there is no CPU fallback.
"""
from .rigid import Rigid, Rotation  # noqa: F401

__all__ = ["Rigid", "Rotation"]
